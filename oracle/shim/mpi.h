/* Single-rank stand-in for <mpi.h>, used ONLY to compile the unmodified reference
 * sources from /root/reference into oracle/_ref (test infrastructure, never shipped).
 * The reference touches this exact symbol set: MPI_Init/Finalize/Comm_size/Comm_rank/
 * Bcast/Send/Recv (main.cpp:50-52,73,99,1439-1583; mpi_util.h:331,355).  With one rank
 * a broadcast is the identity and point-to-point calls are never reached. */
#ifndef PCRAMP_ORACLE_MPI_SHIM_H
#define PCRAMP_ORACLE_MPI_SHIM_H

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)
#define MPI_BYTE 1
#define MPI_UNSIGNED 2

static inline int MPI_Init(int *, char ***) { return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 1; }

#endif
