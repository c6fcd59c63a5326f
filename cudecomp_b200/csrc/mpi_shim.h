// Glue between the MPI_Comm handles of include/mpi_shim/mpi.h and cdb::Comm.
#ifndef CUDECOMP_B200_MPI_SHIM_INTERNAL_H
#define CUDECOMP_B200_MPI_SHIM_INTERNAL_H

#include <mpi.h>

#include "bootstrap.h"

namespace cdb {
CommPtr commFromHandle(int handle); // nullptr if unknown
int registerComm(const CommPtr& c);
} // namespace cdb

#endif
