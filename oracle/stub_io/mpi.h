/* Minimal stand-in for <mpi.h> that lets the reference's io.cpp compile.  TEST INFRASTRUCTURE ONLY.
 * io.cpp uses exactly two MPI names, both on a fatal path (io.cpp:269,272). */
#ifndef ORACLE_STUB_IO_MPI_H
#define ORACLE_STUB_IO_MPI_H
#include <stdlib.h>
#define MPI_COMM_WORLD 0
static inline int MPI_Abort(int comm, int code) { (void)comm; (void)code; abort(); return 0; }
#endif
