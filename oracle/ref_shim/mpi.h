/* empty stand-in: the operator dispatchers only #include <mpi.h>, they never call it */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
#endif
