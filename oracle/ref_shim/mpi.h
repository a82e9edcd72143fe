/* stand-in: the operator sources only #include <mpi.h>; tet4_mass.cpp also times itself with MPI_Wtime */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H
static inline double MPI_Wtime(void) { return 0.0; }
#endif
