// IcpFast::Align on the device (registrators/icp_fast.cc:455-529), double precision.
//
// One ICP iteration = three phases, each a kernel here (the same __device__ phase bodies
// are meant to be driven by a persistent cooperative kernel as well):
//   A  transform + epsilon-approximate 1-NN in the libnabo-compatible tree (icp_fast.cc:
//      486-493, 169-180) + a 2048-bin histogram of the squared distances;
//   B  locate the histogram bin that holds the floor(N*0.7f)-th order statistic
//      (icp_fast.cc:65-90); every match strictly below that bin is accumulated into the
//      point-to-plane normal equations right away (icp_fast.cc:256-302), matches inside
//      the bin are compacted (deterministically, ascending index) as candidates;
//   C  one CTA: exact radix-select of the limit among the candidates, add the candidates
//      <= limit, fixed-order reduction of the per-block partial sums, 6x6 solve, pose
//      update, convergence test and score (icp_fast.cc:204-254, 307-321, 377-405, 513-527).
// Nothing returns to the host between iterations; kernels early-out on state->done.
#include "common.cuh"
#include "kernels.h"
#include "icp_dev.cuh"
#include "linalg_dev.cuh"

namespace smb {
using namespace dev;
namespace {

// ------------------------------------------------------------------------------ prologue
__global__ void __launch_bounds__(256)
mean_partial_kernel(const double* __restrict__ raw, int64_t stride, int n,
                    double* __restrict__ partials) {
  __shared__ double sm[3][8];
  double s[3] = {0.0, 0.0, 0.0};
  const int base = blockIdx.x * 1024;
  for (int r = 0; r < 4; ++r) {
    const int i = base + r * 256 + threadIdx.x;
    if (i < n) { s[0] += raw[i]; s[1] += raw[stride + i]; s[2] += raw[2 * stride + i]; }
  }
  for (int d = 0; d < 3; ++d) {
    double v = s[d];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[d][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sm[threadIdx.x][w];
    partials[blockIdx.x * 4 + threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256)
center_kernel(const double* __restrict__ raw, double* __restrict__ out, int64_t stride, int n,
              const double* __restrict__ partials, int nparts, IcpState* __restrict__ st) {
  __shared__ double mean[3];
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int b = 0; b < nparts; ++b) v += partials[b * 4 + threadIdx.x];
    v = v / (double)n;
    mean[threadIdx.x] = v;
    if (blockIdx.x == 0) st->mean[threadIdx.x] = v;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    out[i] = dsub(raw[i], mean[0]);
    out[stride + i] = dsub(raw[stride + i], mean[1]);
    out[2 * stride + i] = dsub(raw[2 * stride + i], mean[2]);
  }
}

__global__ void fill_buckets_kernel(const double* __restrict__ coord, int64_t cstride,
                                    const double* __restrict__ nrm, int64_t nstride,
                                    const uint32_t* __restrict__ leaf_order, int n,
                                    BucketPoint* __restrict__ bpts, BucketNormal* __restrict__ bnrm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t id = leaf_order[s];
  BucketPoint p;
  p.x = coord[id]; p.y = coord[cstride + id]; p.z = coord[2 * cstride + id]; p.id = id;
  bpts[s] = p;
  if (nrm != nullptr) {
    BucketNormal q;
    q.x = nrm[id]; q.y = nrm[nstride + id]; q.z = nrm[2 * nstride + id]; q.pad = 0.0;
    bnrm[s] = q;
  }
}

// one block: G0 = T_mean^-1 * guess, state reset, histogram clear (icp_fast.cc:460-480)
__global__ void icp_init_kernel(IcpState* __restrict__ st, const double* __restrict__ guess,
                                uint32_t* __restrict__ hist, uint32_t* __restrict__ claim, int nchunks) {
  for (int i = threadIdx.x; i < 2 * kHistBins + 2; i += blockDim.x) hist[i] = 0;  // hist + hist2 + steal cursor + started blocks
  for (int i = threadIdx.x; i < nchunks; i += blockDim.x) claim[i] = 0;            // stamps start at 1
  if (threadIdx.x != 0) return;
  double Tm[16], Tmi[16];
  for (int i = 0; i < 16; ++i) { Tm[i] = (i % 5 == 0) ? 1.0 : 0.0; Tmi[i] = Tm[i]; }
  for (int r = 0; r < 3; ++r) { Tm[12 + r] = st->mean[r]; Tmi[12 + r] = -st->mean[r]; }
  double g[16];
  for (int i = 0; i < 16; ++i) g[i] = guess[i];
  double G0[16];
  la::mul4(Tmi, g, G0);
  for (int i = 0; i < 16; ++i) {
    st->T_mean[i] = Tm[i]; st->G0[i] = G0[i];
    st->T_iter[i] = (i % 5 == 0) ? 1.0 : 0.0;
    st->result[i] = 0.0;
  }
  st->quat_hist[0][0] = 1.0; st->quat_hist[0][1] = 0.0; st->quat_hist[0][2] = 0.0; st->quat_hist[0][3] = 0.0;
  st->trans_hist[0][0] = 0.0; st->trans_hist[0][1] = 0.0; st->trans_hist[0][2] = 0.0;
  st->hist_len = 1; st->iteration = 0; st->done = 0; st->status = 0; st->solve_path = 0;
  st->final_score = 0.0; st->limit = 0.0; st->kept = 0;
}

// 30-bit Morton key on a 0.25 m lattice (coordinates wrap every 256 m, harmless for a key
// that only has to make neighbouring threads spatially close)
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
  v &= 1023u;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// init_source = G0 (x) source (icp_fast.cc:469-471) + the spatial sort key of each point
__global__ void apply_g0_kernel(const double* __restrict__ in, double* __restrict__ out,
                                int64_t stride, int n, const IcpState* __restrict__ st,
                                uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px, py, pz;
  transform_point(st->G0, in[i], in[stride + i], in[2 * stride + i], px, py, pz);
  out[i] = px; out[stride + i] = py; out[2 * stride + i] = pz;
  const uint32_t ix = (uint32_t)(int)floor(px * 4.0), iy = (uint32_t)(int)floor(py * 4.0),
                 iz = (uint32_t)(int)floor(pz * 4.0);
  keys[i] = (uint64_t)(spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2));
  vals[i] = (uint32_t)i;
}

__global__ void visits_key_kernel(const uint8_t* __restrict__ visits, int n, uint64_t* __restrict__ keys,
                                  uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = visits[i];
  vals[i] = (uint32_t)i;
}

// queries in Morton order: neighbouring threads walk the same tree lines and hit the same
// buckets.  Only sums are formed over the source, so its order is free.
__global__ void gather_source_kernel(const double* __restrict__ in, double* __restrict__ out,
                                     int64_t stride, int n, const uint32_t* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = perm[i];
  out[i] = in[s]; out[stride + i] = in[stride + s]; out[2 * stride + i] = in[2 * stride + s];
}

// -------------------------------------------------------------------------------- phase A
__global__ void __launch_bounds__(kKnnThreads)
icp_knn_static_kernel(IcpBuffers b, IcpParams p) {
  __shared__ double T[16];
  if (b.state->done) return;
  // measured: keeping the first 4 pending subtrees per thread in shared memory is SLOWER
  // (2.34 vs 2.00 ms per 30 launches): the 37 KB per block come out of the L1 that caches the
  // tree lines.  The stack therefore stays in (L1-cached) local memory.
  if (threadIdx.x < 16) T[threadIdx.x] = b.state->T_iter[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * kKnnThreads + threadIdx.x;
  if (i < p.n_source) {
    double px, py, pz;
    transform_point(T, b.src0[i], b.src0[b.sstride + i], b.src0[2 * b.sstride + i], px, py, pz);
    int slot; double d2;
    if (p.debug_knn_mode == 0 || p.debug_knn_mode >= 10) {
      int rounds = 0;
      knn1(b.nodes, b.bpts, px, py, pz, p.max_error2, slot, d2,
           p.debug_knn_mode >= 10 ? p.debug_knn_mode - 10 : (1 << 30), &rounds);
      if (b.visits) b.visits[i] = (uint8_t)min(rounds, 255);
    } else {   // profiling aid: truncated variants (results are NOT the k-NN)
      slot = 0; d2 = px * px + py * py + pz * pz + 1.0;
      if (p.debug_knn_mode >= 2) {
        int idx = 0;
        KdNode nd = load_node(b.nodes, 0);
        while (nd.dim != 3) {
          const double q = nd.dim == 0 ? px : (nd.dim == 1 ? py : pz);
          idx = child_idx(idx, (q > nd.cut) ? 1 : 0);
          nd = load_node(b.nodes, idx);
        }
        slot = (int)(__double_as_longlong(nd.cut) & 0xffffffffll);
        if (p.debug_knn_mode >= 3) {
          double head = __longlong_as_double(0x7ff0000000000000ll);
          scan_leaf(b.bpts, nd, px, py, pz, head, slot);
          d2 = head;
        }
      }
    }
    b.slot[i] = slot;
    b.d2[i] = d2;
    // fire-and-forget reduction straight into the 2048-bin global histogram (L2-resident);
    // no block-level staging, so a warp retires as soon as its own queries are done
    if (finite_d2(d2)) atomicAdd(&b.hist[dist_bin(d2)], 1u);
  }
}

// Phase A with work sharing.  The search of one query is the same sequence of passes as in knn1 /
// visit_subtree (pass 0: read-only descent + first bucket; then one pass per far visit), but a
// lane whose query is finished does not idle until the slowest lane of its warp is done: the warp
// claims further 32-query chunks (Morton-consecutive, so still coherent) and hands their queries
// to lanes as they become free.  Chunk c belongs to warp c ("home"); a warp claims its home chunk
// with one atomicExch on the chunk's own stamp (no shared address), and steals from the END of
// the chunk list through a cursor.  With one alignment in flight every warp is resident at once,
// claims its home chunk and finds nothing to steal: the kernel behaves like the static one.  With
// many alignments in flight the warps of late blocks start late, early warps take over their
// chunks, late blocks find them claimed and retire at once: lanes stay busy and a warp slot is
// released as soon as the work runs out (profiles/knn_profile.py: 14 of 32 lanes are active on
// average in the static kernel).
// MEASURED: slower than the static kernel in both regimes — 2.26 vs 1.87 ms
// of k-NN per alignment with one in flight (58 vs 40 registers, the claim round trips, the
// per-pass ballots) and 777 vs 1050 alignments/s with 16 in flight (lanes of one warp end up in
// distant tree regions, every pass waits for more distinct lines).  Option knn_refill, default off.
// Results do not depend on who runs a query: each query's search is self-contained.
constexpr int kRefillMin = 8;   // free lanes before a warp looks for more work

__global__ void __launch_bounds__(kKnnThreads)
icp_knn_kernel(IcpBuffers b, IcpParams p) {
  __shared__ double T[16];
  if (b.state->done) return;
  if (threadIdx.x < 16) T[threadIdx.x] = b.state->T_iter[threadIdx.x];
  const uint32_t stamp = (uint32_t)b.state->iteration + 1u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const int n = p.n_source;
  const int nchunks = (n + 31) >> 5;
  const int home = (blockIdx.x * kKnnThreads + threadIdx.x) >> 5;
  uint32_t* cursor = b.hist + 2 * kHistBins;
  uint32_t* started = cursor + 1;        // blocks of this launch that have begun
  if (threadIdx.x == 0) atomicAdd(started, 1u);
  const KdNode* __restrict__ nodes = b.nodes;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const double me2 = p.max_error2;
  // the warp's window of claimed, not yet started queries (warp-uniform)
  int win_next = 0, win_end = 0;
  bool tried_home = false, exhausted = false;
  // per-lane search state
  bool active = false;
  int qi = 0, phase = 0, best = -1, idx = 0, sp = 0, rounds = 0;
  double qx = 0, qy = 0, qz = 0, head = inf, rd = 0, ox = 0, oy = 0, oz = 0, min_rd = inf;
  StackEntry stack[kMaxStack];
  while (true) {
    // ---- hand queries to free lanes --------------------------------------------------------
    uint32_t freem = __ballot_sync(0xffffffffu, !active);
    while (freem != 0u) {
      if (win_next >= win_end) {
        if (exhausted || (__popc(freem) < kRefillMin && freem != 0xffffffffu)) break;
        int c = -1;
        if (lane == 0) {
          if (!tried_home && home < nchunks && atomicExch(&b.knn_claim[home], stamp) != stamp) {
            c = home;
          } else {
            // steal from the end of the list, but never from a block that has already started: its
            // own warps are claiming those chunks (with one alignment in flight every block has
            // started and this ends after one look)
            while (true) {
              const uint32_t k = atomicAdd(cursor, 1u);
              const int cand = nchunks - 1 - (int)k;
              const int s = (int)*reinterpret_cast<volatile uint32_t*>(started);
              if (cand < 0 || cand < s * (kKnnThreads / 32)) { c = -2; break; }
              if (atomicExch(&b.knn_claim[cand], stamp) != stamp) { c = cand; break; }
            }
          }
        }
        tried_home = true;
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c < 0) { exhausted = true; break; }
        win_next = c << 5;
        win_end = min(win_next + 32, n);
      }
      const int avail = win_end - win_next;
      const int rank = __popc(freem & lt);
      if (!active && rank < avail) {
        qi = win_next + rank;
        transform_point(T, b.src0[qi], b.src0[b.sstride + qi], b.src0[2 * b.sstride + qi], qx, qy, qz);
        active = true; phase = 0; best = -1; idx = 0; sp = 0; rounds = 0;
        head = inf; rd = 0.0; ox = 0.0; oy = 0.0; oz = 0.0; min_rd = inf;
      }
      win_next += min(__popc(freem), avail);
      freem = __ballot_sync(0xffffffffu, !active);
    }
    if (!__any_sync(0xffffffffu, active)) break;
    // ---- one pass: descend to a bucket, scan it, pick the next pending subtree ---------------
    if (active) {
      KdNode nd = load_node(nodes, idx);
      int guard = 0;
      while (nd.dim != 3 && ++guard < 64) {
        const int cd = nd.dim;
        const double q = cd == 0 ? qx : (cd == 1 ? qy : qz);
        const double old_off = cd == 0 ? ox : (cd == 1 ? oy : oz);
        const int right = q > nd.cut ? 1 : 0;
        const int next = child_idx(idx, right);
        const KdNode nd_next = load_node(nodes, next);
        const double new_off = dsub(q, nd.cut);
        const double rd_new = dadd(rd, dadd(-dmul(old_off, old_off), dmul(new_off, new_off)));
        if (phase == 0) {
          min_rd = fmin(min_rd, rd_new);     // == new_off^2 on the first descent (rd = 0, offsets 0)
        } else if (dmul(rd_new, me2) < head && sp < kMaxStack) {
          StackEntry e;
          e.rd = rd_new;
          e.ox = cd == 0 ? new_off : ox;
          e.oy = cd == 1 ? new_off : oy;
          e.oz = cd == 2 ? new_off : oz;
          e.idx = child_idx(idx, 1 - right);
          stack[sp++] = e;
        }
        idx = next;
        nd = nd_next;
      }
      if (nd.dim == 3) scan_leaf(b.bpts, nd, qx, qy, qz, head, best);
      bool fin;
      if (phase == 0) {
        // no far subtree can qualify if even the closest cut plane fails the test (see knn1)
        fin = !(dmul(min_rd, me2) < head);
        phase = 1; idx = 0; rd = 0.0; ox = 0.0; oy = 0.0; oz = 0.0; sp = 0;   // else: replay from the root
      } else {
        ++rounds;
        fin = true;
        while (sp > 0) {
          const StackEntry e = stack[--sp];
          if (dmul(e.rd, me2) < head) {
            idx = e.idx; rd = e.rd; ox = e.ox; oy = e.oy; oz = e.oz;
            fin = false;
            break;
          }
        }
      }
      if (fin) {
        b.slot[qi] = best;
        b.d2[qi] = head;
        if (b.visits) b.visits[qi] = (uint8_t)min(rounds, 255);
        if (finite_d2(head)) atomicAdd(&b.hist[dist_bin(head)], 1u);
        active = false;
      }
    }
  }
}

// Diagnostics (sm_debug_knn_profile): the phase-A search with per-thread clocks.
__global__ void __launch_bounds__(kKnnThreads)
icp_knn_profile_kernel(IcpBuffers b, IcpParams p, int identity, uint32_t* __restrict__ cycles,
                       uint8_t* __restrict__ rounds_out, uint8_t* __restrict__ smid_out,
                       unsigned long long* __restrict__ t0_out, unsigned long long* __restrict__ t1_out) {
  __shared__ double T[16];
  if (threadIdx.x < 16) T[threadIdx.x] = identity ? ((threadIdx.x % 5 == 0) ? 1.0 : 0.0) : b.state->T_iter[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * kKnnThreads + threadIdx.x;
  if (i >= p.n_source) return;
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long c0 = clock64();
  double px, py, pz;
  transform_point(T, b.src0[i], b.src0[b.sstride + i], b.src0[2 * b.sstride + i], px, py, pz);
  int slot, rounds = 0; double d2;
  knn1(b.nodes, b.bpts, px, py, pz, p.max_error2, slot, d2, 1 << 30, &rounds);
  const long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  uint32_t sm;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  cycles[i] = (uint32_t)(c1 - c0) + (slot < 0 ? 0u : 0u) + (d2 < 0.0 ? 1u : 0u);
  rounds_out[i] = (uint8_t)min(rounds, 255);
  smid_out[i] = (uint8_t)sm;
  t0_out[i] = g0; t1_out[i] = g1;
}

// -------------------------------------------------------------------------------- phase B
__global__ void __launch_bounds__(kAccThreads)
icp_accum_kernel(IcpBuffers b, IcpParams p) {
  __shared__ uint32_t warp_tot[8];
  __shared__ BinSel sel_sm;
  __shared__ double T[16];
  __shared__ double red[kAccThreads / 32][kNumSums];
  __shared__ uint32_t cand_warp[kAccItems][kAccThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int tile0 = blockIdx.x * kAccTile;
  // (1) everything that does not depend on the quantile bin is requested up front, for all the
  //     points of the thread at once (these reads overlap the histogram read of select_bin)
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double d2[kAccItems], sx[kAccItems], sy[kAccItems], sz[kAccItems];
  int slot[kAccItems];
#pragma unroll
  for (int r = 0; r < kAccItems; ++r) {
    const int i = tile0 + r * kAccThreads + threadIdx.x;
    const bool in = i < p.n_source;
    d2[r] = in ? b.d2[i] : inf;
    slot[r] = in ? b.slot[i] : -1;
    sx[r] = in ? b.src0[i] : 0.0;
    sy[r] = in ? b.src0[b.sstride + i] : 0.0;
    sz[r] = in ? b.src0[2 * b.sstride + i] : 0.0;
  }
  if (b.state->done) return;
  if (threadIdx.x < 16) T[threadIdx.x] = b.state->T_iter[threadIdx.x];
  const BinSel sel = select_bin(b.hist, p.dist_outlier_ratio, warp_tot, &sel_sm);
  // (2) the matched target point and normal of every point at or below the quantile bin; the
  //     other lanes read entry 0 and are masked out, so the gathers of all items go out together
  bool use[kAccItems], is_cand[kAccItems];
  int bin[kAccItems];
  double2 qxy[kAccItems], nxy[kAccItems];
  double qz[kAccItems], nz[kAccItems];
#pragma unroll
  for (int r = 0; r < kAccItems; ++r) {
    bin[r] = finite_d2(d2[r]) ? dist_bin(d2[r]) : kHistBins;
    use[r] = sel.bin >= 0 && slot[r] >= 0 && bin[r] <= sel.bin;
    const int s = use[r] ? slot[r] : 0;
    qxy[r] = __ldg(reinterpret_cast<const double2*>(b.bpts + s));
    qz[r] = __ldg(reinterpret_cast<const double*>(b.bpts + s) + 2);
    nxy[r] = __ldg(reinterpret_cast<const double2*>(b.bnrm + s));
    nz[r] = __ldg(reinterpret_cast<const double*>(b.bnrm + s) + 2);
  }
  double acc[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) acc[k] = 0.0;
  double F[kAccItems][6], dot[kAccItems], sq[kAccItems];
#pragma unroll
  for (int r = 0; r < kAccItems; ++r) {
    double px, py, pz;
    transform_point(T, sx[r], sy[r], sz[r], px, py, pz);
    BucketPoint q; BucketNormal n;
    q.x = qxy[r].x; q.y = qxy[r].y; q.z = qz[r];
    n.x = nxy[r].x; n.y = nxy[r].y; n.z = nz[r];
    match_terms(px, py, pz, q, n, F[r], dot[r]);
    sq[r] = sqrt(use[r] ? d2[r] : 0.0);
    add_terms_if(acc, F[r], dot[r], sq[r], use[r] && bin[r] < sel.bin);
    is_cand[r] = use[r] && bin[r] == sel.bin;
    if (is_cand[r]) atomicAdd(&b.hist2[sub_bin(d2[r])], 1u);
  }
  // (3) ordered compaction of the quantile-bin members: (item, warp, lane) == ascending point index
  uint32_t m[kAccItems];
#pragma unroll
  for (int r = 0; r < kAccItems; ++r) {
    m[r] = __ballot_sync(0xffffffffu, is_cand[r]);
    if (lane == 0) cand_warp[r][w] = __popc(m[r]);
  }
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int r = 0; r < kAccItems; ++r) {
    uint32_t off = base, tot = 0;
#pragma unroll
    for (int ww = 0; ww < kAccThreads / 32; ++ww) { const uint32_t c = cand_warp[r][ww]; if (ww < w) off += c; tot += c; }
    if (is_cand[r]) {
      const int64_t dst = tile0 + off + __popc(m[r] & ((1u << lane) - 1u));
      b.cand_key[dst] = (unsigned long long)__double_as_longlong(d2[r]);
      double2* o = reinterpret_cast<double2*>(b.cand_terms + 8 * dst);
      o[0] = make_double2(F[r][0], F[r][1]); o[1] = make_double2(F[r][2], F[r][3]);
      o[2] = make_double2(F[r][4], F[r][5]); o[3] = make_double2(dot[r], sq[r]);
    }
    base += tot;
  }
  if (threadIdx.x == 0) b.cand_cnt[blockIdx.x] = base;
  block_reduce_sums<kAccThreads>(acc, red, b.partials + (int64_t)blockIdx.x * 32);
}

__global__ void __launch_bounds__(256)
knn_query_kernel(const KdNode* __restrict__ nodes, const BucketPoint* __restrict__ bpts,
                 const double* __restrict__ query, int64_t qstride, int nq, double max_error2,
                 int tree_levels, int32_t* __restrict__ ids, double* __restrict__ d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int slot; double d;
  knn1(nodes, bpts, query[i], query[qstride + i], query[2 * qstride + i], max_error2, slot, d);
  ids[i] = slot >= 0 ? (int32_t)bpts[slot].id : -1;
  d2[i] = d;
}

}  // namespace

int icp_accum_blocks(int n_source) { return ceil_div(n_source, kAccTile); }

int kd_fill_buckets(const double* coord, int64_t cstride, const double* nrm, int64_t nstride,
                    const uint32_t* leaf_order, int n, BucketPoint* bpts, BucketNormal* bnrm,
                    cudaStream_t stream) {
  fill_buckets_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(coord, cstride, nrm, nstride,
                                                            leaf_order, n, bpts, bnrm);
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

int knn_configure() { return 0; }

int knn_query(const KdNode* nodes, const BucketPoint* bpts, const double* query, int64_t qstride,
              int nq, double max_error2, int tree_levels, int32_t* ids, double* d2,
              cudaStream_t stream) {
  if (nq <= 0) return 0;
  knn_query_kernel<<<ceil_div(nq, 256), 256, 0, stream>>>(
      nodes, bpts, query, qstride, nq, max_error2, tree_levels, ids, d2);
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

// icp_fast.cc:456-480: centre the target, rebuild the tree, G0, initial source transform.
int icp_prologue(const IcpBuffers& b, const IcpParams& p, const double* guess_dev,
                 KdWorkspace& ws, cudaStream_t stream) {
  const int nt = p.n_target, ns = p.n_source;
  const int nparts = ceil_div(nt, 1024);
  mean_partial_kernel<<<nparts, 256, 0, stream>>>(b.tgt_raw, b.tstride, nt, b.mean_partials);
  center_kernel<<<ceil_div(nt, 256), 256, 0, stream>>>(b.tgt_raw, b.tgt, b.tstride, nt,
                                                      b.mean_partials, nparts, b.state);
  int rc = kd_build(b.tgt, b.tstride, nt, 8, ws, b.nodes, b.leaf_order, stream);
  if (rc) return rc;
  rc = kd_fill_buckets(b.tgt, b.tstride, b.nrm, b.tstride, b.leaf_order, nt, b.bpts, b.bnrm, stream);
  if (rc) return rc;
  icp_init_kernel<<<1, 256, 0, stream>>>(b.state, guess_dev, b.hist, b.knn_claim, ceil_div(p.n_source, 32));
  apply_g0_kernel<<<ceil_div(ns, 256), 256, 0, stream>>>(b.src_raw, b.src_g0, b.sstride, ns, b.state,
                                                        b.src_keys[0], b.src_vals[0]);
  rc = radix_sort_pairs_u64(b.src_keys[0], b.src_vals[0], b.src_keys[1], b.src_vals[1], ns, 1,
                            b.sstride, b.src_scratch, stream, 4);
  if (rc) return rc;
  gather_source_kernel<<<ceil_div(ns, 256), 256, 0, stream>>>(b.src_g0, b.src0, b.sstride, ns,
                                                             b.src_vals[0]);
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

// Iterations start_iteration .. start_iteration+count-1 of one Align.  The k-NN of iteration 0
// records how many buckets every query visited; right after it the queries are re-ordered by
// that count (stable: Morton order survives inside a class), so that the lanes of a warp need
// about the same number of sequential traversal rounds in all later iterations (the pose moves
// little between iterations).  Only sums are formed over the source, so its order is free.
int icp_enqueue_iterations(const IcpBuffers& b_in, const IcpParams& p, int start_iteration, int count,
                           cudaStream_t stream, cudaEvent_t* events) {
  const int nb = icp_accum_blocks(p.n_source);
  const bool resort = p.resort_by_visits != 0;
  IcpBuffers b = b_in;
  IcpBuffers b_sorted = b_in;       // what iterations >= 1 read once the re-ordering happened
  b_sorted.src0 = b_in.src_g0;
  b_sorted.visits = nullptr;
  if (!resort) b.visits = nullptr;
  for (int it = 0; it < count; ++it) {
    const int global_it = start_iteration + it;
    const IcpBuffers& bb = (resort && global_it >= 1) ? b_sorted : b;
    if (events) cudaEventRecord(events[4 * it + 0], stream);
    if (p.knn_refill && p.debug_knn_mode == 0)
      icp_knn_kernel<<<ceil_div(p.n_source, kKnnThreads), kKnnThreads, 0, stream>>>(bb, p);
    else
      icp_knn_static_kernel<<<ceil_div(p.n_source, kKnnThreads), kKnnThreads, 0, stream>>>(bb, p);
    if (events) cudaEventRecord(events[4 * it + 1], stream);
    icp_accum_kernel<<<nb, kAccThreads, 0, stream>>>(bb, p);
    if (events) cudaEventRecord(events[4 * it + 2], stream);
    icp_finish_launch(bb, p, nb, stream);
    if (resort && global_it == 0) {
      visits_key_kernel<<<ceil_div(p.n_source, 256), 256, 0, stream>>>(b.visits, p.n_source, b.src_keys[0], b.src_vals[0]);
      int rc = radix_sort_pairs_u64(b.src_keys[0], b.src_vals[0], b.src_keys[1], b.src_vals[1], p.n_source, 1,
                                    b.sstride, b.src_scratch, stream, 1);   // 1 pass: result in [1]
      if (rc) return rc;
      gather_source_kernel<<<ceil_div(p.n_source, 256), 256, 0, stream>>>(b.src0, b.src_g0, b.sstride, p.n_source,
                                                                         b.src_vals[1]);
    }
    if (events) cudaEventRecord(events[4 * it + 3], stream);
  }
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

int icp_knn_profile(const IcpBuffers& b, const IcpParams& p, int identity, uint32_t* cycles, uint8_t* rounds,
                    uint8_t* smid, unsigned long long* t0, unsigned long long* t1, cudaStream_t stream) {
  icp_knn_profile_kernel<<<ceil_div(p.n_source, kKnnThreads), kKnnThreads, 0, stream>>>(b, p, identity, cycles, rounds,
                                                                                       smid, t0, t1);
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace smb
