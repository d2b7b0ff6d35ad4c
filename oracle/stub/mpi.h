/* Empty stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * /root/reference/src/edmd.hpp:27 includes "mpi.h" but edmd.cpp uses no MPI
 * symbol, so an empty header is enough to compile the reference's physics
 * files (edmd.cpp, tetrad.cpp, array.cpp, qcprot.c) where they lie.  No MPI
 * exists in this image; master.cpp / worker.cpp / io.cpp are NOT compiled. */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
#endif
