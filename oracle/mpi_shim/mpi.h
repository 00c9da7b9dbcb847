/*
 * mpi.h -- TEST/BENCH INFRASTRUCTURE ONLY.  A minimal single-node stand-in for MPI so that the reference's own
 * parallel.c (-DSPMD -DMPI, compiled unchanged from /root/reference/src) can run across all host cores on a box that
 * has no MPI installation.  Only what parallel.c reaches at run time is implemented (mpi_shim.c):
 * MPI_Init (forks MOLDY_MPI_NP-1 children), Comm_size/rank, Allreduce/Reduce {INT,FLOAT,DOUBLE} x {SUM,MAX}, Bcast,
 * Abort, Finalize.  The derived-datatype / Allgather calls of par_collect_all (src/parallel.c:842-868) are compiled
 * but only reached under -DMPPMANY; they abort here.
 */
#ifndef MOLDY_MPI_SHIM_H
#define MOLDY_MPI_SHIM_H
#include <stddef.h>

typedef int  MPI_Comm;
typedef int  MPI_Datatype;
typedef int  MPI_Op;
typedef long MPI_Aint;

#define MPI_COMM_WORLD 0
#define MPI_INT    1
#define MPI_FLOAT  2
#define MPI_DOUBLE 3
#define MPI_BYTE   4
#define MPI_UB     5
#define MPI_SUM    1
#define MPI_MAX    2
#define MPI_SUCCESS 0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Allreduce(void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm comm);
int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype old, MPI_Datatype *newt);
int MPI_Type_struct(int count, int *blens, MPI_Aint *displs, MPI_Datatype *types, MPI_Datatype *newt);
int MPI_Type_commit(MPI_Datatype *t);
int MPI_Type_free(MPI_Datatype *t);
int MPI_Allgather(void *send, int ns, MPI_Datatype st, void *recv, int nr, MPI_Datatype rt, MPI_Comm comm);
#endif
