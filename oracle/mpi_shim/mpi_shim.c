/*
 * mpi_shim.c -- TEST/BENCH INFRASTRUCTURE ONLY (see mpi.h).  Ranks are forked processes of one node; collectives go
 * through one anonymous shared mapping: a process-shared pthread barrier and one CHUNK-byte slot per rank.
 * Reductions add the slots in rank order, so every rank gets identical bits (as a real MPI_Allreduce must for
 * Moldy's DESYNC check, src/main.c:262-273).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>
#include "mpi.h"

#define CHUNK (1 << 20)

typedef struct {
   pthread_barrier_t bar;
   volatile int abort_code;
} shm_hdr;

static shm_hdr *hdr;
static char *slots;
static int rank_ = 0, size_ = 1;
static pid_t *kids;

static size_t tsize(MPI_Datatype t) { return t == MPI_DOUBLE ? 8 : t == MPI_BYTE ? 1 : 4; }
static char *slot(int r) { return slots + (size_t)r * CHUNK; }
static void barrier(void) { if (size_ > 1) pthread_barrier_wait(&hdr->bar); }

int MPI_Init(int *argc, char ***argv)
{
   const char *s = getenv("MOLDY_MPI_NP");
   (void)argc; (void)argv;
   size_ = s ? atoi(s) : 1;
   if (size_ < 1) size_ = 1;
   if (size_ == 1) return MPI_SUCCESS;
   const size_t bytes = sizeof(shm_hdr) + 64 + (size_t)size_ * CHUNK;
   char *base = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
   if (base == MAP_FAILED) { perror("mpi_shim: mmap"); exit(3); }
   hdr = (shm_hdr *)base;
   slots = base + ((sizeof(shm_hdr) + 63) & ~(size_t)63);
   pthread_barrierattr_t at;
   pthread_barrierattr_init(&at);
   pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
   pthread_barrier_init(&hdr->bar, &at, (unsigned)size_);
   hdr->abort_code = 0;
   kids = calloc((size_t)size_, sizeof(pid_t));
   fflush(stdout); fflush(stderr);
   for (int r = 1; r < size_; r++) {
      pid_t p = fork();
      if (p < 0) { perror("mpi_shim: fork"); exit(3); }
      if (p == 0) { rank_ = r; free(kids); kids = NULL; return MPI_SUCCESS; }
      kids[r] = p;
   }
   return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
   fflush(stdout); fflush(stderr);
   barrier();
   if (rank_ == 0 && kids)
      for (int r = 1; r < size_; r++) waitpid(kids[r], NULL, 0);
   return MPI_SUCCESS;
}

int MPI_Abort(MPI_Comm comm, int code)
{
   (void)comm;
   fflush(stdout); fflush(stderr);
   if (rank_ == 0 && kids) {
      for (int r = 1; r < size_; r++) kill(kids[r], SIGKILL);
   } else if (size_ > 1) {
      kill(getppid(), SIGTERM);
   }
   _exit(code ? code : 3);
}

int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = size_; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = rank_; return MPI_SUCCESS; }

static void combine(void *dst, int n, MPI_Datatype t, MPI_Op op)
{
   for (int i = 0; i < n; i++) {
      if (t == MPI_DOUBLE) {
         double a = ((double *)slot(0))[i];
         for (int r = 1; r < size_; r++) { double b = ((double *)slot(r))[i]; a = op == MPI_SUM ? a + b : (b > a ? b : a); }
         ((double *)dst)[i] = a;
      } else if (t == MPI_FLOAT) {
         float a = ((float *)slot(0))[i];
         for (int r = 1; r < size_; r++) { float b = ((float *)slot(r))[i]; a = op == MPI_SUM ? a + b : (b > a ? b : a); }
         ((float *)dst)[i] = a;
      } else {
         int a = ((int *)slot(0))[i];
         for (int r = 1; r < size_; r++) { int b = ((int *)slot(r))[i]; a = op == MPI_SUM ? a + b : (b > a ? b : a); }
         ((int *)dst)[i] = a;
      }
   }
}

int MPI_Allreduce(void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, MPI_Comm comm)
{
   (void)comm;
   const size_t ts = tsize(t);
   if (size_ == 1) { memmove(recv, send, ts * (size_t)n); return MPI_SUCCESS; }
   const int per = (int)(CHUNK / ts);
   for (int o = 0; o < n; o += per) {
      const int m = n - o < per ? n - o : per;
      memcpy(slot(rank_), (char *)send + ts * (size_t)o, ts * (size_t)m);
      barrier();
      combine((char *)recv + ts * (size_t)o, m, t, op);
      barrier();
   }
   return MPI_SUCCESS;
}

int MPI_Reduce(void *send, void *recv, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm comm)
{
   if (rank_ == root) return MPI_Allreduce(send, recv, n, t, op, comm);
   void *tmp = malloc(tsize(t) * (size_t)(n > 0 ? n : 1));
   MPI_Allreduce(send, tmp, n, t, op, comm);
   free(tmp);
   return MPI_SUCCESS;
}

int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm comm)
{
   (void)comm;
   const size_t total = tsize(t) * (size_t)n;
   if (size_ == 1) return MPI_SUCCESS;
   for (size_t o = 0; o < total; o += CHUNK) {
      const size_t m = total - o < CHUNK ? total - o : CHUNK;
      if (rank_ == root) memcpy(slot(0), (char *)buf + o, m);
      barrier();
      if (rank_ != root) memcpy((char *)buf + o, slot(0), m);
      barrier();
   }
   return MPI_SUCCESS;
}

static int unsupported(const char *what)
{
   fprintf(stderr, "mpi_shim: %s is not implemented (only reached under -DMPPMANY)\n", what);
   return MPI_Abort(MPI_COMM_WORLD, 3);
}
int MPI_Type_vector(int c, int b, int s, MPI_Datatype o, MPI_Datatype *n) { (void)c; (void)b; (void)s; (void)o; (void)n; return unsupported("MPI_Type_vector"); }
int MPI_Type_struct(int c, int *b, MPI_Aint *d, MPI_Datatype *t, MPI_Datatype *n) { (void)c; (void)b; (void)d; (void)t; (void)n; return unsupported("MPI_Type_struct"); }
int MPI_Type_commit(MPI_Datatype *t) { (void)t; return unsupported("MPI_Type_commit"); }
int MPI_Type_free(MPI_Datatype *t) { (void)t; return unsupported("MPI_Type_free"); }
int MPI_Allgather(void *s, int ns, MPI_Datatype st, void *r, int nr, MPI_Datatype rt, MPI_Comm c)
{ (void)s; (void)ns; (void)st; (void)r; (void)nr; (void)rt; (void)c; return unsupported("MPI_Allgather"); }
