/* Empty stand-in so that the reference's tests/cc/*.cc (which include internal/checks.h, which includes
 * <cutensor.h>) compile in an image without cuTENSOR. Only the names checks.h mentions inside macros that the
 * test executables never expand are declared. Written for this repo; not derived from the cuTENSOR headers. */
#ifndef CUDECOMP_B200_CUTENSOR_STUB_H
#define CUDECOMP_B200_CUTENSOR_STUB_H
typedef int cutensorStatus_t;
#define CUTENSOR_STATUS_SUCCESS 0
static inline const char* cutensorGetErrorString(cutensorStatus_t) { return "cutensor stub"; }
#endif
