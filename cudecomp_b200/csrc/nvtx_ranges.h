// NVTX ranges under the names the reference uses (reference include/internal/nvtx.h, transpose.h:907-953,
// halo.h:317-350, comm_routines.h:260-262), so that timelines captured with Nsight Systems read the same for both
// libraries. NVTX v3 is header-only: without a profiler attached a range costs one indirect call.
#ifndef CUDECOMP_B200_NVTX_RANGES_H
#define CUDECOMP_B200_NVTX_RANGES_H

#include <nvtx3/nvToolsExt.h>

#include <functional>
#include <string>

namespace cdb {

class NvtxRange {
public:
  explicit NvtxRange(const std::string& name) {
    static constexpr uint32_t colors[8] = {0x3366CC, 0xDC3912, 0xFF9900, 0x109618, 0x990099, 0x3B3EAC, 0x0099C6, 0xDD4477};
    nvtxEventAttributes_t ev = {};
    ev.version = NVTX_VERSION;
    ev.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
    ev.colorType = NVTX_COLOR_ARGB;
    ev.color = colors[std::hash<std::string>{}(name) % 8];
    ev.messageType = NVTX_MESSAGE_TYPE_ASCII;
    ev.message.ascii = name.c_str();
    nvtxRangePushEx(&ev);
  }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

} // namespace cdb

#endif
