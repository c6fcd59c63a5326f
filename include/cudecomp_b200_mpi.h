/*
 * cudecomp_b200_mpi.h -- adapter for applications that run on a REAL MPI installation.
 *
 * The reference's cudecompInit takes the application's MPI_Comm (reference include/cudecomp.h:249) and uses MPI for its
 * control plane. This engine has its own control plane (a TCP mesh between the ranks; pencil data only ever moves GPU
 * to GPU over NVLink), so all it needs from MPI is "who am I, how many are we, where does rank 0 listen". This header
 * obtains exactly that through the application's own MPI and hands it to the library through an entry point that
 * carries no MPI type, so ONE library binary serves every MPI implementation (MPI_Comm is an int in MPICH and a pointer
 * in Open MPI).
 *
 * Usage -- no change to the application's source:
 *   mpicc  -Iinclude -include cudecomp_b200_mpi.h app.c -Lcudecomp_b200/lib -lcudecomp_realmpi ...
 * i.e. compile against the site's <mpi.h> (do NOT add include/mpi_shim to the include path), force-include this header
 * (or include it after cudecomp.h), and link libcudecomp_realmpi.so: the same code as libcudecomp.so, linked with a
 * version script that exports only cudecomp* symbols, so that the MPI-subset shim inside the library (for machines
 * without MPI) can never shadow the real MPI_* functions.
 *
 * cudecompInit(&handle, comm) then resolves to cudecompB200InitFromMPI below. `comm` may be any intra-communicator; all
 * of its ranks must make the call (it is collective, like the reference's). One communicator per process lifetime: the
 * library keeps one control-plane mesh. Fortran applications (cudecompInit_F, reference src/cudecomp_m.cuf) need the
 * module's C shim compiled with this header in the same way.
 */
#ifndef CUDECOMP_B200_MPI_H
#define CUDECOMP_B200_MPI_H

#include <mpi.h>

#ifdef CUDECOMP_B200_MPI_SHIM
#error "cudecomp_b200_mpi.h is for a real MPI: remove include/mpi_shim from the include path"
#endif

#include <string.h>

#include "cudecomp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Exported by libcudecomp*.so (also declared in cudecomp_b200_ext.h). */
cudecompResult_t cudecompB200InitBootstrap(cudecompHandle_t* handle, int32_t rank, int32_t nranks, const char* root_addr,
                                           int32_t root_port);
cudecompResult_t cudecompB200PickBootstrapPort(int32_t* port);

#ifdef __cplusplus
}
#endif

static inline cudecompResult_t cudecompB200InitFromMPI(cudecompHandle_t* handle, MPI_Comm comm) {
  int rank = -1, size = 0;
  /* [0..255] host name of rank 0, then its port as text: one broadcast */
  char rendezvous[256 + 16];
  if (MPI_Comm_rank(comm, &rank) != MPI_SUCCESS || MPI_Comm_size(comm, &size) != MPI_SUCCESS)
    return CUDECOMP_RESULT_MPI_ERROR;
  memset(rendezvous, 0, sizeof(rendezvous));
  if (rank == 0) {
    int32_t port = 0;
    int len = 0;
    char name[MPI_MAX_PROCESSOR_NAME + 1];
    cudecompResult_t res = cudecompB200PickBootstrapPort(&port);
    if (res != CUDECOMP_RESULT_SUCCESS) port = -1; /* still take part in the broadcast */
    memset(name, 0, sizeof(name));
    if (size == 1 || MPI_Get_processor_name(name, &len) != MPI_SUCCESS || len <= 0 || len > 255)
      strcpy(name, "127.0.0.1");
    strcpy(rendezvous, name);
    {
      /* decimal port without pulling in <stdio.h> */
      char digits[16];
      int n = 0, p = port, i;
      if (p <= 0) digits[n++] = '0';
      while (p > 0 && n < 15) {
        digits[n++] = (char)('0' + p % 10);
        p /= 10;
      }
      for (i = 0; i < n; ++i) rendezvous[256 + i] = digits[n - 1 - i];
    }
  }
  if (MPI_Bcast(rendezvous, (int)sizeof(rendezvous), MPI_CHAR, 0, comm) != MPI_SUCCESS) return CUDECOMP_RESULT_MPI_ERROR;
  {
    int32_t port = 0;
    int i;
    for (i = 256; i < 256 + 15 && rendezvous[i] >= '0' && rendezvous[i] <= '9'; ++i) port = port * 10 + (rendezvous[i] - '0');
    if (port <= 0) return CUDECOMP_RESULT_INTERNAL_ERROR;
    rendezvous[255] = 0;
    {
      /* ranks on rank 0's own host dial the loopback address: host names do not always resolve (containers) */
      char mine[MPI_MAX_PROCESSOR_NAME + 1];
      int len = 0;
      memset(mine, 0, sizeof(mine));
      if (MPI_Get_processor_name(mine, &len) == MPI_SUCCESS && len > 0 && strcmp(mine, rendezvous) == 0)
        strcpy(rendezvous, "127.0.0.1");
    }
    return cudecompB200InitBootstrap(handle, rank, size, rendezvous, port);
  }
}

/* the application's cudecompInit(&handle, comm) calls now take the adapter */
#define cudecompInit(handle, comm) cudecompB200InitFromMPI((handle), (comm))

#endif
