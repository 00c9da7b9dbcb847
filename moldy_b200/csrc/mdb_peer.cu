// mdb_peer.cu -- the replicated-data multi-GPU layer: par_rsum/par_dsum (src/parallel.c:549-588, call sites
// src/accel.c:531-535) as kernels of our own over NVLink/NVSwitch peer memory.
//
// Every rank (one GPU) owns a WINDOW of device memory that all other ranks map -- directly when the ranks are engines
// of one process (mdb_peer_connect, cudaDeviceEnablePeerAccess), through CUDA IPC handles when they are processes
// (mdb_peer_handle / mdb_peer_open; torchrun ranks of bench.py).  The same kernels serve both:
//
//   k_peer_barrier  flag barrier: store the epoch into every peer's flag word (st.release.sys), spin on our own
//                   (ld.acquire.sys).  No host round trip, no NCCL launch; ~3 us on NVSwitch.
//   k_peer_sum      dst[i] = sum_p window_p[i] in rank order over up to four index ranges: the structure-factor
//                   all-reduce (every rank sums all of it: one-shot, 1.2 MB at 10^6 sites), and the force
//                   reduce-scatter (a rank sums only its own slice of the sites, plus the 16 scalars).  Fixed order:
//                   every rank gets identical bits (what Moldy's DESYNC check relies on, src/main.c:262-273).
//   k_peer_gather   pull the slices the other ranks own out of their windows: all-gather of the site co-ordinates a
//                   rank uploaded over its own PCIe link, of the reduced forces, of the c-of-m/quaternion slices.
//
// A step is cut into phases so that one host thread can drive all ranks of a process (enqueue phase A on every
// device, then the barrier on every device, ...) without a blocking call ever waiting for a kernel whose partner
// has not been enqueued yet; separate processes simply run the phases in sequence.
//   A  zero out[parity]; cell build; real-space sum over this rank's batches; structure-factor partial sums of this
//      rank's charged sites -> window.psum                                       (src/force.c:856, moldy.tex:3441-3466)
//   -- barrier
//   B  psum_total = sum_p window_p.psum; energy/stress (rank 0) and k-space forces on own sites -> out[parity]
//   -- barrier
//   C  window.red[own slice] = sum_p window_p.out[parity][own slice]; scalars summed by every rank
//   (-- barrier, D: all-gather of red, when every rank needs every force: the reference's par_rsum semantics)
// out[] is double-buffered by step parity, so a rank may start step n+1 while a peer still reads step n's block; every
// other window region is protected by the two barriers of the following step.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "mdb_internal.h"

static constexpr int MAXP = MDB_MAX_PEERS;
static constexpr size_t OFF_FLAGS = 0, FLAG_BYTES = 1024;      // u32 flags[MAXP], then u32 error word at [MAXP]

struct PeerWin { unsigned char *base[MAXP]; };
struct PeerRanges { long long start[4], len[4]; int n; };
struct PeerBounds { long long lo[MAXP + 1]; };

struct mdb_peer {
   mdb_engine *e = nullptr;
   int rank = 0, world = 1;
   bool ipc = false;
   unsigned char *win = nullptr;
   size_t bytes = 0;
   PeerWin W{};
   bool opened[MAXP] = {};
   size_t n = 0, nslots_cap = 0, in_cap = 0;
   size_t off_psum = 0, off_out[2] = {0, 0}, off_xyz = 0, off_in = 0, off_red = 0;
   unsigned epoch = 0;
   int parity = 0;
   double *d_psum_tot = nullptr;
   PeerBounds site_b{}, in_b{};
   double *h_pin = nullptr; size_t pin_cap = 0;         // pinned staging of the host-facing calls
   long barriers = 0;
   // the structure-factor pass does not need the cell lists: it runs on a side stream beside the (latency-bound) cell
   // build and sub-list compaction and shares the SMs with the pair kernel afterwards
   cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
   asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
   unsigned v;
   asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
   unsigned long long t;
   asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
   return t;
}

// One warp; lane q signals peer q and waits for peer q's signal.  A peer that never arrives (a rank that died) trips
// the 20 s time-out: the error word is set and the kernel returns, so the GPU is never left spinning.
__global__ void __launch_bounds__(32) k_peer_barrier(PeerWin W, int rank, int world, unsigned epoch)
{
   const int q = threadIdx.x;
   if (q >= world) return;
   __threadfence_system();
   st_release_sys(reinterpret_cast<unsigned *>(W.base[q] + OFF_FLAGS) + rank, epoch);
   const unsigned *mine = reinterpret_cast<const unsigned *>(W.base[rank] + OFF_FLAGS) + q;
   const unsigned long long t0 = globaltimer_ns();
   while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (globaltimer_ns() - t0 > 20000000000ULL) {
         reinterpret_cast<unsigned *>(W.base[rank] + OFF_FLAGS)[MAXP] = 1u;
         break;
      }
   }
   __threadfence_system();
}

// dst[start_s - dst_base + i] = sum over ranks of window[src_off][start_s + i], ranks in order 0..world-1.
// Peer lines are read with ld.global.cg (L2 of the owner, never our L1).
__global__ void __launch_bounds__(256) k_peer_sum(PeerWin W, int world, size_t src_off, double *__restrict__ dst,
                                                  long long dst_base, PeerRanges R)
{
   long long total = 0;
   for (int s = 0; s < R.n; s++) total += R.len[s];
   for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
      long long i = t;
      int s = 0;
      while (i >= R.len[s]) { i -= R.len[s]; s++; }
      const long long idx = R.start[s] + i;
      // all peers' values are requested before the first add: `world` NVLink round trips in flight per thread instead of
      // one after the other; the sum itself stays in rank order (identical bits on every rank)
      double v[MAXP];
#pragma unroll
      for (int p = 0; p < MAXP; p++)
         v[p] = p < world ? __ldcg(reinterpret_cast<const double *>(W.base[p] + src_off) + idx) : 0.0;
      double acc = 0.0;
#pragma unroll
      for (int p = 0; p < MAXP; p++)
         if (p < world) acc += v[p];
      dst[idx - dst_base] = acc;
   }
}

// All-gather by pulling: for `rows` rows of `row_len` doubles at `off` of every window, copy the part of each row that
// peer p owns (B.lo[p] .. B.lo[p+1]) from p's window into ours.
__global__ void __launch_bounds__(256) k_peer_gather(PeerWin W, int rank, int world, size_t off, int rows, long long row_len,
                                                     PeerBounds B)
{
   const long long total = (long long)rows * row_len;
   double *mine = reinterpret_cast<double *>(W.base[rank] + off);
   for (long long t = blockIdx.x * 256LL + threadIdx.x; t < total; t += (long long)gridDim.x * 256) {
      const long long s = t % row_len;
      int p = 0;
      while (p + 1 < world && s >= B.lo[p + 1]) p++;
      if (p != rank) mine[t] = __ldcg(reinterpret_cast<const double *>(W.base[p] + off) + t);
   }
}

static int grid_for(long long n) { return (int)std::min<long long>(148 * 8, std::max<long long>(1, (n + 255) / 256)); }

// ---------------------------------------------------------------------------------------------------------------
extern "C" mdb_peer *mdb_peer_create(mdb_engine *e, int rank, int world)
{
   if (!e || !e->configured) { mdb_set_error("mdb_peer_create: engine not configured"); return nullptr; }
   if (world < 1 || world > MAXP || rank < 0 || rank >= world) { mdb_set_error("mdb_peer_create: bad rank/world"); return nullptr; }
   if (cudaSetDevice(e->device) != cudaSuccess) { mdb_set_error("mdb_peer_create: cudaSetDevice"); return nullptr; }
   mdb_peer *p = new mdb_peer();
   p->e = e; p->rank = rank; p->world = world;
   p->n = (size_t)e->cfg.nsites;
   p->nslots_cap = (size_t)std::max(e->T.nslots, 1) * 3 / 2 + 64;       // head-room for a breathing cell
   p->in_cap = 7 * p->n;                                                 // c-of-m + quaternions of <= n molecules
   const size_t outd = mdb_out_doubles((int)p->n);
   size_t off = FLAG_BYTES;
   p->off_psum = off;   off = align_up(off + sizeof(double) * 8 * p->nslots_cap);
   p->off_out[0] = off; off = align_up(off + sizeof(double) * outd);
   p->off_out[1] = off; off = align_up(off + sizeof(double) * outd);
   p->off_xyz = off;    off = align_up(off + sizeof(double) * 3 * p->n);
   p->off_in = off;     off = align_up(off + sizeof(double) * p->in_cap);
   p->off_red = off;    off = align_up(off + sizeof(double) * outd);
   p->bytes = off;
   if (cudaMalloc(&p->win, p->bytes) != cudaSuccess || cudaMemset(p->win, 0, p->bytes) != cudaSuccess ||
       cudaMalloc(&p->d_psum_tot, sizeof(double) * 8 * p->nslots_cap) != cudaSuccess) {
      mdb_set_error("mdb_peer_create: out of device memory");
      if (p->win) cudaFree(p->win);
      delete p;
      return nullptr;
   }
   p->W.base[rank] = p->win;
   for (int r = 0; r <= world; r++) {
      p->site_b.lo[r] = (long long)p->n * r / world;
      p->in_b.lo[r] = 0;
   }
   mdb_set_partition(e, rank, world);
   double *xyz = reinterpret_cast<double *>(p->win + p->off_xyz);
   mdb_set_sites_device(e, xyz, xyz + p->n, xyz + 2 * p->n, nullptr);
   e->sites_set = false;
   return p;
}

extern "C" void mdb_peer_destroy(mdb_peer *p)
{
   if (!p) return;
   cudaSetDevice(p->e->device);
   cudaDeviceSynchronize();
   for (int r = 0; r < p->world; r++)
      if (p->ipc && p->opened[r]) cudaIpcCloseMemHandle(p->W.base[r]);
   if (p->win) cudaFree(p->win);
   if (p->side) cudaStreamDestroy(p->side);
   if (p->ev_fork) cudaEventDestroy(p->ev_fork);
   if (p->ev_join) cudaEventDestroy(p->ev_join);
   if (p->d_psum_tot) cudaFree(p->d_psum_tot);
   if (p->h_pin) cudaFreeHost(p->h_pin);
   delete p;
}

extern "C" size_t mdb_peer_window_bytes(const mdb_peer *p) { return p->bytes; }

// ranks are processes: a 64-byte CUDA IPC handle of the window per rank, exchanged by the caller (torch.distributed)
extern "C" int mdb_peer_handle(mdb_peer *p, void *handle)
{
   static_assert(sizeof(cudaIpcMemHandle_t) == MDB_PEER_HANDLE_BYTES, "IPC handle size");
   MDB_CUDA(cudaSetDevice(p->e->device));
   cudaIpcMemHandle_t h;
   MDB_CUDA(cudaIpcGetMemHandle(&h, p->win));
   memcpy(handle, &h, sizeof h);
   return 0;
}
extern "C" int mdb_peer_open(mdb_peer *p, const void *handles)
{
   MDB_CUDA(cudaSetDevice(p->e->device));
   for (int r = 0; r < p->world; r++) {
      if (r == p->rank) continue;
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char *)handles + (size_t)r * MDB_PEER_HANDLE_BYTES, sizeof h);
      void *ptr = nullptr;
      MDB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      p->W.base[r] = (unsigned char *)ptr;
      p->opened[r] = true;
   }
   p->ipc = true;
   return 0;
}
// ranks are engines of this process: map the windows directly
extern "C" int mdb_peer_connect(mdb_peer *const *peers, int world)
{
   for (int a = 0; a < world; a++) {
      mdb_peer *p = peers[a];
      if (p->world != world || p->rank != a) { mdb_set_error("mdb_peer_connect: peers must be passed in rank order"); return -1; }
      MDB_CUDA(cudaSetDevice(p->e->device));
      for (int b = 0; b < world; b++) {
         if (b == a) continue;
         if (peers[b]->e->device != p->e->device) {
            int can = 0;
            MDB_CUDA(cudaDeviceCanAccessPeer(&can, p->e->device, peers[b]->e->device));
            if (!can) { mdb_set_error("mdb_peer_connect: devices cannot access each other's memory (no NVLink/P2P)"); return -1; }
            cudaError_t rc = cudaDeviceEnablePeerAccess(peers[b]->e->device, 0);
            if (rc != cudaSuccess && rc != cudaErrorPeerAccessAlreadyEnabled) {
               mdb_set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(rc));
               return -1;
            }
            cudaGetLastError();
         }
         p->W.base[b] = peers[b]->win;
      }
   }
   return 0;
}

// Ownership of the result: rank r reduces sites [bounds[r], bounds[r+1]) (original site order; world+1 ascending values,
// bounds[0] = 0, bounds[world] = nsites).  Default: equal shares.  eval_forces() cuts at molecule boundaries.
extern "C" int mdb_peer_set_site_bounds(mdb_peer *p, const long long *bounds)
{
   for (int r = 0; r <= p->world; r++) {
      if (bounds[r] < 0 || bounds[r] > (long long)p->n || (r && bounds[r] < bounds[r - 1])) { mdb_set_error("mdb_peer_set_site_bounds"); return -1; }
      p->site_b.lo[r] = bounds[r];
   }
   return 0;
}

extern "C" int mdb_peer_barrier(mdb_peer *p, void *stream)
{
   if (p->world == 1) return 0;
   MDB_CUDA(cudaSetDevice(p->e->device));
   p->epoch++;
   k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(p->W, p->rank, p->world, p->epoch);
   p->e->launches++; p->barriers++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// 1 when a barrier of this rank timed out since the last call (synchronises `stream`)
extern "C" int mdb_peer_error(mdb_peer *p, void *stream)
{
   unsigned v = 0;
   if (cudaMemcpyAsync(&v, p->win + OFF_FLAGS + 4 * MAXP, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess ||
       cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
      return -1;
   return (int)v;
}

// ---- inputs ------------------------------------------------------------------------------------------------------
// This rank's slice of three HOST rows (full-length arrays, only [bounds[rank], bounds[rank+1]) is read) -> window.
extern "C" int mdb_peer_sites_host_slice(mdb_peer *p, const double *x, const double *y, const double *z, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(p->e->device));
   const long long lo = p->site_b.lo[p->rank], hi = p->site_b.lo[p->rank + 1];
   double *xyz = reinterpret_cast<double *>(p->win + p->off_xyz);
   const double *rows[3] = {x, y, z};
   for (int a = 0; a < 3 && hi > lo; a++)
      MDB_CUDA(cudaMemcpyAsync(xyz + a * p->n + lo, rows[a] + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyHostToDevice, st));
   return 0;
}
// All of the three rows (a rank that holds the full configuration, e.g. at start-up)
extern "C" int mdb_peer_sites_host_all(mdb_peer *p, const double *x, const double *y, const double *z, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(p->e->device));
   double *xyz = reinterpret_cast<double *>(p->win + p->off_xyz);
   MDB_CUDA(cudaMemcpyAsync(xyz, x, sizeof(double) * p->n, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(xyz + p->n, y, sizeof(double) * p->n, cudaMemcpyHostToDevice, st));
   MDB_CUDA(cudaMemcpyAsync(xyz + 2 * p->n, z, sizeof(double) * p->n, cudaMemcpyHostToDevice, st));
   mdb_set_sites_device(p->e, xyz, xyz + p->n, xyz + 2 * p->n, nullptr);
   return 0;
}
// after a barrier: pull the other ranks' slices; the engine then holds all sites
extern "C" int mdb_peer_sites_gather(mdb_peer *p, void *stream)
{
   MDB_CUDA(cudaSetDevice(p->e->device));
   if (p->world > 1) {
      k_peer_gather<<<grid_for(3 * (long long)p->n), 256, 0, (cudaStream_t)stream>>>(p->W, p->rank, p->world, p->off_xyz, 3,
                                                                                   (long long)p->n, p->site_b);
      p->e->launches++;
      MDB_CUDA(cudaGetLastError());
   }
   double *xyz = reinterpret_cast<double *>(p->win + p->off_xyz);
   mdb_set_sites_device(p->e, xyz, xyz + p->n, xyz + 2 * p->n, nullptr);      // (eval_forces switches the engine to its own rows)
   return 0;
}
extern "C" double *mdb_peer_sites(mdb_peer *p) { return reinterpret_cast<double *>(p->win + p->off_xyz); }

// generic input block (eval_forces: [c-of-m 3 nmols | quaternions 4 nmols_q]) of `len` doubles: the rank uploads the
// part [len r/P, len (r+1)/P) from HOST memory, the rest is pulled from the peers after a barrier
extern "C" double *mdb_peer_in(mdb_peer *p) { return reinterpret_cast<double *>(p->win + p->off_in); }
extern "C" int mdb_peer_in_host_slice(mdb_peer *p, const double *h_in, size_t len, void *stream)
{
   if (len > p->in_cap) { mdb_set_error("mdb_peer_in_host_slice: block larger than the window"); return -1; }
   MDB_CUDA(cudaSetDevice(p->e->device));
   for (int r = 0; r <= p->world; r++) p->in_b.lo[r] = (long long)len * r / p->world;
   const long long lo = p->in_b.lo[p->rank], hi = p->in_b.lo[p->rank + 1];
   if (hi > lo)
      MDB_CUDA(cudaMemcpyAsync(mdb_peer_in(p) + lo, h_in + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyHostToDevice,
                               (cudaStream_t)stream));
   return 0;
}
extern "C" int mdb_peer_in_gather(mdb_peer *p, size_t len, void *stream)
{
   if (p->world == 1) return 0;
   MDB_CUDA(cudaSetDevice(p->e->device));
   k_peer_gather<<<grid_for((long long)len), 256, 0, (cudaStream_t)stream>>>(p->W, p->rank, p->world, p->off_in, 1, (long long)len,
                                                                           p->in_b);
   p->e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// all-gather of ONE row of `row_len` doubles at `off_doubles` of the input block, owned in the pieces bounds[r] .. bounds[r+1]
// (the resident NVE step of a device group: every rank moves its own molecules and pulls the others')
extern "C" int mdb_peer_in_gather_bounds(mdb_peer *p, size_t off_doubles, long long row_len, const long long *bounds, void *stream)
{
   if (p->world == 1 || row_len <= 0) return 0;
   if (off_doubles + (size_t)row_len > p->in_cap) { mdb_set_error("mdb_peer_in_gather_bounds: row outside the window"); return -1; }
   MDB_CUDA(cudaSetDevice(p->e->device));
   PeerBounds B{};
   for (int r = 0; r <= p->world; r++) B.lo[r] = bounds[r];
   k_peer_gather<<<grid_for(row_len), 256, 0, (cudaStream_t)stream>>>(p->W, p->rank, p->world, p->off_in + sizeof(double) * off_doubles, 1,
                                                                      row_len, B);
   p->e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// ---- the step ----------------------------------------------------------------------------------------------------
static double *out_block(mdb_peer *p) { return reinterpret_cast<double *>(p->win + p->off_out[p->parity]); }

// phase A.  what: bit 0 real space, bit 1 reciprocal space; bit 2: continue the phase begun by an earlier call
// (no new result block, no cell build) -- lets bench.py bracket cells / pair / k-space with events.
extern "C" int mdb_peer_phase_a(mdb_peer *p, int what, void *stream)
{
   mdb_engine *e = p->e;
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(e->device));
   if ((size_t)e->cfg.nsites != p->n || (size_t)std::max(e->T.nslots, 1) > p->nslots_cap) {
      mdb_set_error("mdb_peer: the engine was reconfigured beyond the window's capacity; recreate the peer");
      return -1;
   }
   mdb_set_partition(e, p->rank, p->world);
   if (!(what & 4)) p->parity ^= 1;
   double *out = out_block(p);
   const bool recip = (what & 2) && e->cfg.do_recip;
   double *psum = p->world == 1 ? p->d_psum_tot : reinterpret_cast<double *>(p->win + p->off_psum);
   static const bool no_side = getenv("MDB_PEER_NO_SIDE") != nullptr;
   const bool fork = recip && (what & 1) && !(what & 4) && !no_side;
   if (fork) {
      if (!p->side) {
         int lo = 0, hi = 0;                          // MDB_PEER_SIDE_PRIO=1: the structure-factor blocks go first
         static const bool prio = getenv("MDB_PEER_SIDE_PRIO") && atoi(getenv("MDB_PEER_SIDE_PRIO")) != 0;
         MDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
         MDB_CUDA(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, prio ? hi : lo));
         MDB_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
         MDB_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
      }
      MDB_CUDA(cudaEventRecord(p->ev_fork, st));                   // the sites are complete at this point of `st`
      MDB_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
      if (mdb_launch_recip_partial(e, psum, p->side)) return -1;
      MDB_CUDA(cudaEventRecord(p->ev_join, p->side));
   }
   if (!(what & 4) && (mdb_zero_out(e, out, st) || mdb_build_cells(e, st))) return -1;
   // the pair passes start when the structure-factor pass has ended (MDB_PEER_SHARE=1: beside it, as measured in round 2)
   static const bool share = getenv("MDB_PEER_SHARE") && atoi(getenv("MDB_PEER_SHARE")) != 0;
   if (fork && !share) e->pre_pair_wait = p->ev_join;
   const int rc_real = (what & 1) ? mdb_force_real(e, out, st) : 0;
   e->pre_pair_wait = nullptr;
   if (rc_real) return -1;
   if (fork) MDB_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));
   else if (recip && mdb_launch_recip_partial(e, psum, st)) return -1;
   return 0;
}

// phase B (after a barrier): structure-factor all-reduce + second k-space pass
extern "C" int mdb_peer_phase_b(mdb_peer *p, int what, void *stream)
{
   mdb_engine *e = p->e;
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(e->device));
   if (!(what & 2) || !e->cfg.do_recip) return 0;
   if (p->world > 1) {
      PeerRanges R{};
      R.n = 1; R.start[0] = 0; R.len[0] = 8LL * std::max(e->T.nslots, 1);
      k_peer_sum<<<grid_for(R.len[0]), 256, 0, st>>>(p->W, p->world, p->off_psum, p->d_psum_tot, 0, R);
      e->launches++;
      MDB_CUDA(cudaGetLastError());
   }
   return mdb_launch_recip_finish(e, p->d_psum_tot, out_block(p), st);
}

// phase C (after a barrier): reduce-scatter of the forces (own slice of the sites) + the 16 scalars on every rank
extern "C" int mdb_peer_phase_c(mdb_peer *p, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(p->e->device));
   const long long lo = p->site_b.lo[p->rank], hi = p->site_b.lo[p->rank + 1], n = (long long)p->n;
   PeerRanges R{};
   R.n = 4;
   for (int a = 0; a < 3; a++) { R.start[a] = a * n + lo; R.len[a] = hi - lo; }
   R.start[3] = 3 * n; R.len[3] = MDB_OUT_SCALARS;
   k_peer_sum<<<grid_for(3 * (hi - lo) + MDB_OUT_SCALARS), 256, 0, st>>>(p->W, p->world, p->off_out[p->parity],
                                                                        reinterpret_cast<double *>(p->win + p->off_red), 0, R);
   p->e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// phase D (after a barrier): all-gather of the reduced forces, for callers that need every force on every rank
// (par_rsum's semantics: with phase C this is a two-shot all-reduce)
extern "C" int mdb_peer_phase_d(mdb_peer *p, void *stream)
{
   if (p->world == 1) return 0;
   MDB_CUDA(cudaSetDevice(p->e->device));
   k_peer_gather<<<grid_for(3 * (long long)p->n), 256, 0, (cudaStream_t)stream>>>(p->W, p->rank, p->world, p->off_red, 3,
                                                                                (long long)p->n, p->site_b);
   p->e->launches++;
   MDB_CUDA(cudaGetLastError());
   return 0;
}

// the reduced block [fx(N) | fy(N) | fz(N) | pe, pe_recip | stress[9] | pad] (DEVICE): own slice after phase C,
// everything after phase D
extern "C" double *mdb_peer_result(mdb_peer *p) { return reinterpret_cast<double *>(p->win + p->off_red); }
extern "C" double *mdb_peer_partial(mdb_peer *p) { return out_block(p); }
extern "C" void mdb_peer_slice(const mdb_peer *p, long long lohi[2]) { lohi[0] = p->site_b.lo[p->rank]; lohi[1] = p->site_b.lo[p->rank + 1]; }
extern "C" long mdb_peer_barriers(const mdb_peer *p) { return p->barriers; }

// One process per rank: the whole force evaluation on `stream`.  what: bit 0 real, bit 1 reciprocal space; gather: also
// run phase D.  Nothing synchronises.
extern "C" int mdb_peer_step(mdb_peer *p, int what, int gather, void *stream)
{
   if (mdb_peer_phase_a(p, what, stream) || mdb_peer_barrier(p, stream) || mdb_peer_phase_b(p, what, stream) ||
       mdb_peer_barrier(p, stream) || mdb_peer_phase_c(p, stream))
      return -1;
   if (gather && (mdb_peer_barrier(p, stream) || mdb_peer_phase_d(p, stream))) return -1;
   return 0;
}

// D2H of this rank's slice of the reduced forces into three HOST rows (full-length arrays; only the slice is written)
// and of the 16 scalars; synchronises `stream`.
extern "C" int mdb_peer_read_slice_host(mdb_peer *p, double *fx, double *fy, double *fz, double *scal16, void *stream)
{
   cudaStream_t st = (cudaStream_t)stream;
   MDB_CUDA(cudaSetDevice(p->e->device));
   const long long lo = p->site_b.lo[p->rank], hi = p->site_b.lo[p->rank + 1];
   const double *red = mdb_peer_result(p);
   double *rows[3] = {fx, fy, fz};
   for (int a = 0; a < 3 && hi > lo; a++)
      MDB_CUDA(cudaMemcpyAsync(rows[a] + lo, red + a * p->n + lo, sizeof(double) * (size_t)(hi - lo), cudaMemcpyDeviceToHost, st));
   if (scal16) MDB_CUDA(cudaMemcpyAsync(scal16, red + 3 * p->n, sizeof(double) * MDB_OUT_SCALARS, cudaMemcpyDeviceToHost, st));
   MDB_CUDA(cudaStreamSynchronize(st));
   return 0;
}
