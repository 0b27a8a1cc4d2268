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
#include "linalg_dev.cuh"

namespace smb {
namespace {

constexpr int kKnnThreads = 256;
constexpr int kAccThreads = 256;
constexpr int kAccItems = 2;
constexpr int kAccTile = kAccThreads * kAccItems;  // points per accumulate block
constexpr int kNumSums = 29;                        // 21 (A upper) + 6 (b) + sum sqrt + count
constexpr int kFinThreads = 1024;
constexpr int kMaxStack = 32;

// monotone bin of a non-negative finite double: 1/32-octave resolution over 2^-40..2^24
__device__ __forceinline__ int dist_bin(double d2) {
  const long long bits = __double_as_longlong(d2);
  const int key = (int)(bits >> 47) - ((1023 - 40) << 5);
  return min(max(key, 0), kHistBins - 1);
}

__device__ __forceinline__ bool finite_d2(double d) { return d < __longlong_as_double(0x7ff0000000000000ll); }

// p = T (x) s with the reference's accumulation order (cloud_types.cc:288-296)
__device__ __forceinline__ void transform_point(const double* __restrict__ T, double x, double y,
                                                double z, double& px, double& py, double& pz) {
  px = dadd(dadd(dadd(dmul(T[0], x), dmul(T[4], y)), dmul(T[8], z)), T[12]);
  py = dadd(dadd(dadd(dmul(T[1], x), dmul(T[5], y)), dmul(T[9], z)), T[13]);
  pz = dadd(dadd(dadd(dmul(T[2], x), dmul(T[6], y)), dmul(T[10], z)), T[14]);
}

__device__ __forceinline__ KdNode load_node(const KdNode* __restrict__ nodes, int h) {
  const int4 v = __ldg(reinterpret_cast<const int4*>(nodes + h));
  KdNode n;
  n.cut = __hiloint2double(v.y, v.x);
  n.dim = v.z; n.pad = v.w;
  return n;
}

__device__ __forceinline__ void scan_leaf(const BucketPoint* __restrict__ bpts, const KdNode& leaf,
                                          double qx, double qy, double qz, double& head,
                                          int& best) {
  const long long packed = __double_as_longlong(leaf.cut);
  const int first = (int)(packed & 0xffffffffll), count = (int)(packed >> 32);
  for (int k = 0; k < count; ++k) {
    const double2 xy = __ldg(reinterpret_cast<const double2*>(bpts + first + k));
    const double z = __ldg(reinterpret_cast<const double*>(bpts + first + k) + 2);
    const double dx = dsub(qx, xy.x), dy = dsub(qy, xy.y), dz = dsub(qz, z);
    const double dist = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
    if (dist < head) { head = dist; best = first + k; }  // strict: first visited wins
  }
}

struct StackEntry {
  double rd, ox, oy, oz;
  int h;
};

// libnabo recurseKnn (k=1, allowSelfMatch, maxRadius=inf) made iterative.  A far subtree
// is pushed only if it passes the pruning test against the head known at push time (the
// head can only shrink, so this never drops a subtree the recursion would visit) and is
// re-tested at pop time, which is exactly when the recursion tests it.  A read-only first
// descent seeds the head so the stack stays almost empty for epsilon = 3.16.
__device__ __forceinline__ void knn1(const KdNode* __restrict__ nodes,
                                     const BucketPoint* __restrict__ bpts, double qx, double qy,
                                     double qz, double max_error2, int& best_slot, double& best_d2) {
  double head = __longlong_as_double(0x7ff0000000000000ll);
  int best = -1;
  int h = 0;
  KdNode nd = load_node(nodes, 0);
  while (nd.dim != 3) {
    const double q = nd.dim == 0 ? qx : (nd.dim == 1 ? qy : qz);
    h = 2 * h + 1 + ((dsub(q, nd.cut) > 0.0) ? 1 : 0);
    nd = load_node(nodes, h);
  }
  const int leaf0 = h;
  scan_leaf(bpts, nd, qx, qy, qz, head, best);

  StackEntry stack[kMaxStack];
  int sp = 0;
  double rd = 0.0, ox = 0.0, oy = 0.0, oz = 0.0;
  h = 0;
  while (true) {
    while (true) {
      nd = load_node(nodes, h);
      if (nd.dim == 3) {
        if (h != leaf0) scan_leaf(bpts, nd, qx, qy, qz, head, best);
        break;
      }
      const int cd = nd.dim;
      const double q = cd == 0 ? qx : (cd == 1 ? qy : qz);
      const double old_off = cd == 0 ? ox : (cd == 1 ? oy : oz);
      const double new_off = dsub(q, nd.cut);
      const int right = new_off > 0.0 ? 1 : 0;
      // rd += - old_off*old_off + new_off*new_off
      const double rd_new = dadd(rd, dadd(-dmul(old_off, old_off), dmul(new_off, new_off)));
      if (dmul(rd_new, max_error2) < head && sp < kMaxStack) {
        StackEntry e;
        e.rd = rd_new;
        e.ox = cd == 0 ? new_off : ox;
        e.oy = cd == 1 ? new_off : oy;
        e.oz = cd == 2 ? new_off : oz;
        e.h = 2 * h + 1 + (1 - right);
        stack[sp++] = e;
      }
      h = 2 * h + 1 + right;
    }
    bool found = false;
    while (sp > 0) {
      const StackEntry e = stack[--sp];
      if (dmul(e.rd, max_error2) < head) {
        h = e.h; rd = e.rd; ox = e.ox; oy = e.oy; oz = e.oz;
        found = true;
        break;
      }
    }
    if (!found) break;
  }
  best_slot = best;
  best_d2 = head;
}

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
                                uint32_t* __restrict__ hist) {
  for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) hist[i] = 0;
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

__global__ void apply_g0_kernel(const double* __restrict__ in, double* __restrict__ out,
                                int64_t stride, int n, const IcpState* __restrict__ st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px, py, pz;
  transform_point(st->G0, in[i], in[stride + i], in[2 * stride + i], px, py, pz);
  out[i] = px; out[stride + i] = py; out[2 * stride + i] = pz;
}

// -------------------------------------------------------------------------------- phase A
__global__ void __launch_bounds__(kKnnThreads)
icp_knn_kernel(IcpBuffers b, IcpParams p) {
  __shared__ uint32_t hist[kHistBins];
  __shared__ double T[16];
  if (b.state->done) return;
  for (int i = threadIdx.x; i < kHistBins; i += kKnnThreads) hist[i] = 0;
  if (threadIdx.x < 16) T[threadIdx.x] = b.state->T_iter[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * kKnnThreads + threadIdx.x;
  if (i < p.n_source) {
    double px, py, pz;
    transform_point(T, b.src0[i], b.src0[b.sstride + i], b.src0[2 * b.sstride + i], px, py, pz);
    int slot; double d2;
    knn1(b.nodes, b.bpts, px, py, pz, p.max_error2, slot, d2);
    b.slot[i] = slot;
    b.d2[i] = d2;
    if (finite_d2(d2)) atomicAdd(&hist[dist_bin(d2)], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kHistBins; k += kKnnThreads) {
    const uint32_t c = hist[k];
    if (c) atomicAdd(&b.hist[k], c);
  }
}

// -------------------------------------------------------------------------------- phase B
struct BinSel { int bin; int below; int qi; int nvalid; };

// every block locates the quantile bin from the global histogram (2048 bins, 8 per thread
// on the first 256 threads; all threads of the block must call this)
__device__ __forceinline__ BinSel select_bin(const uint32_t* __restrict__ ghist, float ratio,
                                             uint32_t* warp_tot /*[8]*/, BinSel* out_sm) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const bool active = t < 256;
  uint32_t c[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) { c[k] = active ? ghist[t * 8 + k] : 0u; s += c[k]; }
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (active && lane == 31) warp_tot[w] = incl;
  __syncthreads();
  uint32_t base = 0, total = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) { const uint32_t v = warp_tot[ww]; if (ww < w) base += v; total += v; }
  const uint32_t excl = base + incl - s;
  // icp_fast.cc:82-89: quantile == 1.0 -> max element, else index int(size * quantile)
  const double q = (double)ratio;
  int qi = (q == 1.0) ? (int)total - 1 : (int)((double)total * q);
  if (qi > (int)total - 1) qi = (int)total - 1;
  if (active && total > 0 && (uint32_t)qi >= excl && (uint32_t)qi < excl + s) {
    uint32_t run = excl;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if ((uint32_t)qi >= run && (uint32_t)qi < run + c[k]) {
        out_sm->bin = t * 8 + k; out_sm->below = (int)run;
      }
      run += c[k];
    }
    out_sm->qi = qi; out_sm->nvalid = (int)total;
  }
  if (total == 0 && t == 0) { out_sm->bin = -1; out_sm->below = 0; out_sm->qi = 0; out_sm->nvalid = 0; }
  __syncthreads();
  return *out_sm;
}

// contribution of one match to the normal equations (icp_fast.cc:268-302)
__device__ __forceinline__ void accumulate_match(double* acc, double px, double py, double pz,
                                                 const BucketPoint& q, const BucketNormal& n,
                                                 double d2) {
  double F[6];
  F[0] = py * n.z - pz * n.y;
  F[1] = pz * n.x - px * n.z;
  F[2] = px * n.y - py * n.x;
  F[3] = n.x; F[4] = n.y; F[5] = n.z;
  const double dot = (px - q.x) * n.x + (py - q.y) * n.y + (pz - q.z) * n.z;
  int k = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = r; c < 6; ++c) acc[k++] += F[r] * F[c];
#pragma unroll
  for (int r = 0; r < 6; ++r) acc[21 + r] += F[r] * dot;
  acc[27] += sqrt(d2);
  acc[28] += 1.0;
}

__device__ __forceinline__ void load_match(const IcpBuffers& b, const double* T, int i,
                                           double& px, double& py, double& pz, BucketPoint& q,
                                           BucketNormal& n) {
  transform_point(T, b.src0[i], b.src0[b.sstride + i], b.src0[2 * b.sstride + i], px, py, pz);
  const int s = b.slot[i];
  const double2 qxy = __ldg(reinterpret_cast<const double2*>(b.bpts + s));
  q.x = qxy.x; q.y = qxy.y;
  q.z = __ldg(reinterpret_cast<const double*>(b.bpts + s) + 2);
  const double2 nxy = __ldg(reinterpret_cast<const double2*>(b.bnrm + s));
  n.x = nxy.x; n.y = nxy.y;
  n.z = __ldg(reinterpret_cast<const double*>(b.bnrm + s) + 2);
}

// deterministic block reduction of kNumSums doubles (fixed shuffle tree, fixed warp order)
template <int NT>
__device__ __forceinline__ void block_reduce_sums(double* acc, double (*sm)[kNumSums], double* out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[w][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kNumSums) {
    double v = 0.0;
    for (int ww = 0; ww < NT / 32; ++ww) v += sm[ww][threadIdx.x];
    out[threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kAccThreads)
icp_accum_kernel(IcpBuffers b, IcpParams p) {
  __shared__ uint32_t warp_tot[8];
  __shared__ BinSel sel_sm;
  __shared__ double T[16];
  __shared__ double red[kAccThreads / 32][kNumSums];
  __shared__ uint32_t cand_warp[kAccThreads / 32];
  if (b.state->done) return;
  if (threadIdx.x < 16) T[threadIdx.x] = b.state->T_iter[threadIdx.x];
  const BinSel sel = select_bin(b.hist, p.dist_outlier_ratio, warp_tot, &sel_sm);
  double acc[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) acc[k] = 0.0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t cand_base = 0;
  const int tile0 = blockIdx.x * kAccTile;
  for (int r = 0; r < kAccItems; ++r) {
    const int i = tile0 + r * kAccThreads + threadIdx.x;
    bool is_cand = false;
    if (i < p.n_source && sel.bin >= 0) {
      const double d2 = b.d2[i];
      if (finite_d2(d2)) {
        const int bin = dist_bin(d2);
        if (bin < sel.bin) {
          double px, py, pz; BucketPoint q; BucketNormal n;
          load_match(b, T, i, px, py, pz, q, n);
          accumulate_match(acc, px, py, pz, q, n, d2);
        } else if (bin == sel.bin) {
          is_cand = true;
        }
      }
    }
    // ordered compaction of candidates: (round, warp, lane) == ascending point index
    const uint32_t m = __ballot_sync(0xffffffffu, is_cand);
    if (lane == 0) cand_warp[w] = __popc(m);
    __syncthreads();
    uint32_t off = cand_base, tot = 0;
#pragma unroll
    for (int ww = 0; ww < kAccThreads / 32; ++ww) { const uint32_t c = cand_warp[ww]; if (ww < w) off += c; tot += c; }
    if (is_cand) b.cand_idx[tile0 + off + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    cand_base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) b.cand_cnt[blockIdx.x] = cand_base;
  block_reduce_sums<kAccThreads>(acc, red, b.partials + (int64_t)blockIdx.x * 32);
}

// -------------------------------------------------------------------------------- phase C
__global__ void __launch_bounds__(kFinThreads)
icp_finish_kernel(IcpBuffers b, IcpParams p, int nblocks_b) {
  __shared__ uint32_t warp_tot[8];
  __shared__ BinSel sel_sm;
  __shared__ uint32_t sh_hist[256];
  __shared__ uint32_t sh_scan[kFinThreads];
  __shared__ double red[kFinThreads / 32][kNumSums];
  __shared__ double sums[32];
  __shared__ double cand_sums[32];
  __shared__ double T[16];
  __shared__ unsigned long long sh_prefix;
  __shared__ int sh_rank;
  IcpState* st = b.state;
  if (st->done) return;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t < 16) T[t] = st->T_iter[t];
  const BinSel sel = select_bin(b.hist, p.dist_outlier_ratio, warp_tot, &sel_sm);
  if (sel.nvalid == 0) {
    if (t == 0) { st->status = -2; st->done = 1; }  // CHECK(!values.empty()), icp_fast.cc:81
    return;
  }
  // ---- compact the per-block candidate lists (ascending point index) ---------------------
  uint32_t* cand = b.cand_idx;            // in: per-block regions
  uint32_t* flat = b.cand_idx + ((int64_t)nblocks_b * kAccTile);  // out: flat list
  uint32_t total = 0;
  for (int base = 0; base < nblocks_b; base += kFinThreads) {
    const int blk = base + t;
    const uint32_t c = blk < nblocks_b ? b.cand_cnt[blk] : 0;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) sh_scan[w] = incl;
    __syncthreads();
    uint32_t wb = 0, tot = 0;
    for (int ww = 0; ww < kFinThreads / 32; ++ww) { const uint32_t v = sh_scan[ww]; if (ww < w) wb += v; tot += v; }
    const uint32_t off = total + wb + incl - c;
    for (uint32_t k = 0; k < c; ++k) flat[off + k] = cand[(int64_t)blk * kAccTile + k];
    total += tot;
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  // ---- exact radix select (MSB first, 8 x 8 bits) of rank (qi - below) -------------------
  if (t == 0) { sh_prefix = 0ull; sh_rank = sel.qi - sel.below; }
  __syncthreads();
  for (int pass = 7; pass >= 0; --pass) {
    const int shift = pass * 8;
    if (t < 256) sh_hist[t] = 0;
    __syncthreads();
    const unsigned long long prefix = sh_prefix;
    const unsigned long long mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
    for (uint32_t k = t; k < total; k += kFinThreads) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(b.d2[flat[k]]);
      if ((key & mask) == prefix) atomicAdd(&sh_hist[(key >> shift) & 255ull], 1u);
    }
    __syncthreads();
    if (t == 0) {
      int r = sh_rank;
      int d = 0;
      for (; d < 255; ++d) { const int c = (int)sh_hist[d]; if (r < c) break; r -= c; }
      sh_rank = r;
      sh_prefix = prefix | ((unsigned long long)d << shift);
    }
    __syncthreads();
  }
  const double limit = __longlong_as_double((long long)sh_prefix);
  // ---- candidates with d2 <= limit (icp_fast.cc:497-498), ascending-index strided order ---
  double acc[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) acc[k] = 0.0;
  for (uint32_t k = t; k < total; k += kFinThreads) {
    const int i = (int)flat[k];
    const double d2 = b.d2[i];
    if (d2 <= limit) {
      double px, py, pz; BucketPoint q; BucketNormal n;
      load_match(b, T, i, px, py, pz, q, n);
      accumulate_match(acc, px, py, pz, q, n, d2);
    }
  }
  block_reduce_sums<kFinThreads>(acc, red, cand_sums);
  // ---- fixed-order reduction of the phase-B partials -------------------------------------
  if (w < kNumSums) {
    double v = 0.0;
    for (int blk = lane; blk < nblocks_b; blk += 32) v += b.partials[(int64_t)blk * 32 + w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sums[w] = v;
  }
  // clear the histogram for the next iteration
  for (int k = t; k < kHistBins; k += kFinThreads) b.hist[k] = 0;
  __syncthreads();
  if (t != 0) return;
  // ---- serial tail: solve, pose update, convergence (icp_fast.cc:204-254,307-321,377-405) --
  double S[kNumSums];
  for (int k = 0; k < kNumSums; ++k) S[k] = sums[k] + cand_sums[k];
  const double kept = S[28];
  st->limit = limit;
  st->kept = (long long)kept;
  if (!(kept > 0.0)) { st->status = -3; st->done = 1; return; }  // "no point to minimize"
  double A[36], rhs[6], x[6];
  {
    int k = 0;
    for (int r = 0; r < 6; ++r)
      for (int c = r; c < 6; ++c) { A[r * 6 + c] = S[k]; A[c * 6 + r] = S[k]; ++k; }
    for (int r = 0; r < 6; ++r) rhs[r] = -S[21 + r];
  }
  st->solve_path = la::solve_possibly_underdetermined(A, rhs, x);
  const double sq = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  const double angle = sqrt(sq);
  double axis[3] = {x[0], x[1], x[2]};
  if (sq > 0.0) { const double nr = sqrt(sq); axis[0] = x[0] / nr; axis[1] = x[1] / nr; axis[2] = x[2] / nr; }
  double R[9];
  la::angle_axis_to_rotation(angle, axis, R);
  bool has_nan = false;
  for (int i = 0; i < 9; ++i) has_nan |= isnan(R[i]);
  for (int i = 3; i < 6; ++i) has_nan |= isnan(x[i]);
  if (has_nan) { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  double dT[16], Tn[16];
  for (int i = 0; i < 16; ++i) dT[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) dT[r + 4 * c] = R[r * 3 + c];
    dT[12 + r] = x[3 + r];
  }
  la::mul4(dT, T, Tn);
  for (int i = 0; i < 16; ++i) st->T_iter[i] = Tn[i];
  const int iteration = st->iteration + 1;
  st->iteration = iteration;
  // rotation / translation history ring (only the last 5 entries are ever read)
  double Rm[9], qn[4];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Rm[r * 3 + c] = Tn[r + 4 * c];
  la::rotation_to_quaternion(Rm, qn);
  const int len = st->hist_len;
  for (int k = 0; k < 4; ++k) st->quat_hist[len % 5][k] = qn[k];
  for (int k = 0; k < 3; ++k) st->trans_hist[len % 5][k] = Tn[12 + k];
  st->hist_len = len + 1;
  bool conv = false;
  if (!p.disable_convergence && len + 1 > 4) {
    double rot = 0.0, tr = 0.0;
    for (int i = len; i >= len + 1 - 4; --i) {
      rot += fabs(la::quaternion_angular_distance(st->quat_hist[i % 5], st->quat_hist[(i - 1) % 5]));
      const double dx = st->trans_hist[i % 5][0] - st->trans_hist[(i - 1) % 5][0];
      const double dy = st->trans_hist[i % 5][1] - st->trans_hist[(i - 1) % 5][1];
      const double dz = st->trans_hist[i % 5][2] - st->trans_hist[(i - 1) % 5][2];
      tr += fabs(sqrt(dx * dx + dy * dy + dz * dz));
    }
    rot /= 4.0; tr /= 4.0;
    conv = rot < 0.001 && tr < 0.01;
  }
  if (conv || iteration >= p.max_iteration) {
    st->final_score = exp(-(S[27] / kept));
    double tmp[16], res[16];
    la::mul4(st->T_mean, Tn, tmp);           // (T_mean * T_iter) * G0, icp_fast.cc:527
    la::mul4(tmp, st->G0, res);
    for (int i = 0; i < 16; ++i) st->result[i] = res[i];
    st->done = 1;
  }
}

__global__ void __launch_bounds__(256)
knn_query_kernel(const KdNode* __restrict__ nodes, const BucketPoint* __restrict__ bpts,
                 const double* __restrict__ query, int64_t qstride, int nq, double max_error2,
                 int32_t* __restrict__ ids, double* __restrict__ d2) {
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

int knn_query(const KdNode* nodes, const BucketPoint* bpts, const double* query, int64_t qstride,
              int nq, double max_error2, int32_t* ids, double* d2, cudaStream_t stream) {
  if (nq <= 0) return 0;
  knn_query_kernel<<<ceil_div(nq, 256), 256, 0, stream>>>(nodes, bpts, query, qstride, nq,
                                                          max_error2, ids, d2);
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
  icp_init_kernel<<<1, 256, 0, stream>>>(b.state, guess_dev, b.hist);
  apply_g0_kernel<<<ceil_div(ns, 256), 256, 0, stream>>>(b.src_raw, b.src0, b.sstride, ns, b.state);
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

int icp_enqueue_iterations(const IcpBuffers& b, const IcpParams& p, int count,
                           cudaStream_t stream, cudaEvent_t* events) {
  const int nb = icp_accum_blocks(p.n_source);
  for (int it = 0; it < count; ++it) {
    if (events) cudaEventRecord(events[4 * it + 0], stream);
    icp_knn_kernel<<<ceil_div(p.n_source, kKnnThreads), kKnnThreads, 0, stream>>>(b, p);
    if (events) cudaEventRecord(events[4 * it + 1], stream);
    icp_accum_kernel<<<nb, kAccThreads, 0, stream>>>(b, p);
    if (events) cudaEventRecord(events[4 * it + 2], stream);
    icp_finish_kernel<<<1, kFinThreads, 0, stream>>>(b, p, nb);
    if (events) cudaEventRecord(events[4 * it + 3], stream);
  }
  SMB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace smb
