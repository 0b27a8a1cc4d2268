// Phase C of the ICP iteration (see icp.cu), one CTA:
//   * exact floor(N*0.7f)-th order statistic of the squared match distances
//     (icp_fast.cc:65-90).  Phase A left a 2048-bin histogram, phase B a second-level
//     2048-bin histogram of the members of the quantile bin, so only the handful of keys in
//     ONE second-level bin are ranked exactly;
//   * members of the quantile bin with d2 <= limit join the normal equations, and all
//     per-block partial sums are reduced in a fixed order (deterministic result);
//   * warp 0: 6x6 solve, pose update, convergence test, score (icp_fast.cc:204-254,
//     307-321, 377-405, 513-527).  The code is deliberately compact (rolled loops, shared
//     memory): a single warp fetching kilobytes of straight-line code cold costs far more
//     than the arithmetic.
#include "icp_dev.cuh"
#include "linalg_dev.cuh"

namespace smb {
using namespace dev;
namespace {

constexpr int kSelThreads = 1024;
constexpr int kMaxExactKeys = 1024;

struct SolveSmem {
  double S[32];       // reduced sums
  double A[36];       // normal matrix (row-major, symmetric)
  double W[36];       // Cholesky factor (lower)
  double Li[36];      // inverse of the factor (certificate)
  double rhs[6], y[6], x[6];
};

__device__ __noinline__ int solve_exact_path(const double* A, const double* rhs, double* x) {
  return la::solve_possibly_underdetermined(A, rhs, x);
}

// icp_fast.cc:204-254 with a fast path.  The reference asks Eigen's fullPivHouseholderQr
// whether A is invertible (rank 6 at threshold 6*eps*maxpivot) and then solves with LLT.
// For the SPD normal matrix we first factor A = L L^T and bound its condition number by
// trace(A) * ||L^-1||_F^2 >= lambda_max / lambda_min.  Full-pivoting QR keeps every
// |r_kk| >= sigma_k / (6 * 2^5), so a bound below 1e9 certifies rank 6 with five orders
// of magnitude to spare and the pivoted QR is skipped; otherwise the exact restatement
// (QR rank, min-norm branch, SVD fallback) runs.  Returns the path: 0 LLT, 1 min-norm, 2 SVD.
__device__ __forceinline__ int solve_warp(SolveSmem& sm, int lane) {
#pragma unroll 1
  for (int idx = lane; idx < 36; idx += 32) {
    const int r = idx / 6, c = idx % 6;
    const int lo = min(r, c), hi = max(r, c);
    const double v = sm.S[lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo)];
    sm.A[idx] = v; sm.W[idx] = v; sm.Li[idx] = 0.0;
  }
  if (lane < 6) sm.rhs[lane] = -sm.S[21 + lane];
  __syncwarp();
  bool ok = true;
#pragma unroll 1
  for (int k = 0; k < 6; ++k) {
    const double d = sm.W[k * 6 + k];
    ok = ok && (d > 0.0);
    const double lkk = sqrt(d);
    __syncwarp();
    if (lane > k && lane < 6) sm.W[lane * 6 + k] = sm.W[lane * 6 + k] / lkk;
    if (lane == k) sm.W[k * 6 + k] = lkk;
    __syncwarp();
#pragma unroll 1
    for (int idx = lane; idx < 36; idx += 32) {
      const int i = idx / 6, j = idx % 6;
      if (j > k && i >= j) sm.W[idx] -= sm.W[i * 6 + k] * sm.W[j * 6 + k];
    }
    __syncwarp();
  }
  // certificate: columns of L^-1 by forward substitution, one lane per column
  double fro2 = 0.0;
  if (lane < 6 && ok) {
    const int c = lane;
#pragma unroll 1
    for (int i = c; i < 6; ++i) {
      double v = (i == c) ? 1.0 : 0.0;
#pragma unroll 1
      for (int j = c; j < i; ++j) v -= sm.W[i * 6 + j] * sm.Li[j * 6 + c];
      v /= sm.W[i * 6 + i];
      sm.Li[i * 6 + c] = v;
      fro2 += v * v;
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) fro2 += __shfl_xor_sync(0xffffffffu, fro2, o);
  fro2 = __shfl_sync(0xffffffffu, fro2, 0);
  const double trace = sm.A[0] + sm.A[7] + sm.A[14] + sm.A[21] + sm.A[28] + sm.A[35];
  const double cond_bound = trace * fro2;
  const bool certified = ok && (cond_bound < 1e9);   // false for NaN/inf as well
  int path = 0;
  if (lane == 0) {
    if (certified) {
      // LLT::solve: forward then backward substitution (icp_fast.cc:252)
#pragma unroll 1
      for (int i = 0; i < 6; ++i) {
        double s = sm.rhs[i];
#pragma unroll 1
        for (int j = 0; j < i; ++j) s -= sm.W[i * 6 + j] * sm.y[j];
        sm.y[i] = s / sm.W[i * 6 + i];
      }
#pragma unroll 1
      for (int i = 5; i >= 0; --i) {
        double s = sm.y[i];
#pragma unroll 1
        for (int j = i + 1; j < 6; ++j) s -= sm.W[j * 6 + i] * sm.x[j];
        sm.x[i] = s / sm.W[i * 6 + i];
      }
    } else {
      path = solve_exact_path(sm.A, sm.rhs, sm.x);
    }
  }
  __syncwarp();
  return path;
}

// pose update + convergence + score, lane 0 of warp 0 (icp_fast.cc:307-321,377-405,513-527)
__device__ __forceinline__ void update_pose(IcpState* st, const IcpParams& p, SolveSmem& sm,
                                            const double* T, int path) {
  const double* x = sm.x;
  const double kept = sm.S[28];
  st->solve_path = path;
  const double sq = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  const double angle = sqrt(sq);
  double axis[3] = {x[0], x[1], x[2]};
  if (sq > 0.0) { axis[0] = x[0] / angle; axis[1] = x[1] / angle; axis[2] = x[2] / angle; }
  double R[9];
  la::angle_axis_to_rotation(angle, axis, R);
  bool has_nan = false;
#pragma unroll 1
  for (int i = 0; i < 9; ++i) has_nan |= isnan(R[i]);
#pragma unroll 1
  for (int i = 3; i < 6; ++i) has_nan |= isnan(x[i]);
  if (has_nan) {
#pragma unroll 1
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  double dT[16], Tn[16];
#pragma unroll 1
  for (int i = 0; i < 16; ++i) dT[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll 1
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) dT[r + 4 * c] = R[r * 3 + c];
    dT[12 + r] = x[3 + r];
  }
  la::mul4(dT, T, Tn);
#pragma unroll 1
  for (int i = 0; i < 16; ++i) st->T_iter[i] = Tn[i];
  const int iteration = st->iteration + 1;
  st->iteration = iteration;
  bool conv = false;
  if (!p.disable_convergence) {
    double Rm[9], qn[4];
#pragma unroll 1
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rm[r * 3 + c] = Tn[r + 4 * c];
    la::rotation_to_quaternion(Rm, qn);
    const int len = st->hist_len;
    for (int k = 0; k < 4; ++k) st->quat_hist[len % 5][k] = qn[k];
    for (int k = 0; k < 3; ++k) st->trans_hist[len % 5][k] = Tn[12 + k];
    st->hist_len = len + 1;
    if (len + 1 > 4) {
      double rot = 0.0, tr = 0.0;
#pragma unroll 1
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
  }
  if (conv || iteration >= p.max_iteration) {
    st->final_score = exp(-(sm.S[27] / kept));
    double tmp[16], res[16];
    la::mul4(st->T_mean, Tn, tmp);           // (T_mean * T_iter) * G0, icp_fast.cc:527
    la::mul4(tmp, st->G0, res);
#pragma unroll 1
    for (int i = 0; i < 16; ++i) st->result[i] = res[i];
    st->done = 1;
  }
}

__global__ void __launch_bounds__(kSelThreads)
icp_finish_kernel(IcpBuffers b, IcpParams p, int nblocks_b) {
  __shared__ uint32_t warp_tot[32];
  __shared__ BinSel sel_sm;
  __shared__ int sub_sel[3];                 // second-level bin, count below it, its count
  __shared__ unsigned long long keys[kMaxExactKeys];
  __shared__ uint32_t nkeys;
  __shared__ unsigned long long limit_bits;
  __shared__ uint32_t sh_hist[256];
  __shared__ int sh_rank;
  __shared__ double red[kSelThreads / 32][kNumSums];
  __shared__ double cand_sums[32];
  __shared__ double T[16];
  __shared__ SolveSmem solve;
  IcpState* st = b.state;
  if (st->done) return;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t < 16) T[t] = st->T_iter[t];
  if (t == 0) { nkeys = 0; sub_sel[0] = -1; sub_sel[1] = 0; sub_sel[2] = 0; limit_bits = 0ull; }
  const BinSel sel = select_bin(b.hist, p.dist_outlier_ratio, warp_tot, &sel_sm);
  if (sel.nvalid == 0) {
    if (t == 0) { st->status = -2; st->done = 1; }  // CHECK(!values.empty()), icp_fast.cc:81
    return;
  }
  const int r = sel.qi - sel.below;          // rank inside the quantile bin
  // ---- flatten the per-block candidate lists (ascending point index) ----------------------
  uint32_t* flat_idx = b.cand_idx + (int64_t)nblocks_b * kAccTile;
  unsigned long long* flat_key = b.cand_key + (int64_t)nblocks_b * kAccTile;
  uint32_t total = 0;
#pragma unroll 1
  for (int base = 0; base < nblocks_b; base += kSelThreads) {
    const int blk = base + t;
    const uint32_t c = blk < nblocks_b ? b.cand_cnt[blk] : 0;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __syncthreads();
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint32_t wb = 0, tot = 0;
#pragma unroll 1
    for (int ww = 0; ww < kSelThreads / 32; ++ww) { const uint32_t v = warp_tot[ww]; if (ww < w) wb += v; tot += v; }
    const uint32_t off = total + wb + incl - c;
#pragma unroll 1
    for (uint32_t k = 0; k < c; ++k) {
      flat_idx[off + k] = b.cand_idx[(int64_t)blk * kAccTile + k];
      flat_key[off + k] = b.cand_key[(int64_t)blk * kAccTile + k];
    }
    total += tot;
  }
  __syncthreads();
  bool fallback = clamp_bin(sel.bin);
  if (!fallback) {
    // ---- second-level histogram: 2 bins per thread, block exclusive scan ------------------
    const uint32_t c0 = b.hist2[2 * t], c1 = b.hist2[2 * t + 1], s = c0 + c1;
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint32_t base = 0;
#pragma unroll 1
    for (int ww = 0; ww < w; ++ww) base += warp_tot[ww];
    const uint32_t excl = base + incl - s;
    if ((uint32_t)r >= excl && (uint32_t)r < excl + s) {
      if ((uint32_t)r < excl + c0) { sub_sel[0] = 2 * t; sub_sel[1] = (int)excl; sub_sel[2] = (int)c0; }
      else { sub_sel[0] = 2 * t + 1; sub_sel[1] = (int)(excl + c0); sub_sel[2] = (int)c1; }
    }
    __syncthreads();
    if (sub_sel[0] < 0 || sub_sel[2] > kMaxExactKeys) fallback = true;   // block-uniform
  }
  if (!fallback) {
    // ---- gather the keys of that one bin, rank them by counting ---------------------------
    const int sb = sub_sel[0], r2 = r - sub_sel[1];
#pragma unroll 1
    for (uint32_t k = t; k < total; k += kSelThreads) {
      const unsigned long long key = flat_key[k];
      if ((int)((key >> 36) & 2047ull) == sb) keys[atomicAdd(&nkeys, 1u)] = key;
    }
    __syncthreads();
    const int nk = (int)nkeys;
    if (t < nk) {
      const unsigned long long my = keys[t];
      int less = 0, eq = 0;
#pragma unroll 1
      for (int j = 0; j < nk; ++j) { const unsigned long long o = keys[j]; less += o < my; eq += o == my; }
      if (less <= r2 && r2 < less + eq) limit_bits = my;
    }
    __syncthreads();
  } else {
    // ---- general path (clamp bins, e.g. identical clouds: all d2 == 0): MSB radix select ---
    if (t == 0) { sh_rank = r; limit_bits = 0ull; }
    __syncthreads();
#pragma unroll 1
    for (int pass = 7; pass >= 0; --pass) {
      const int shift = pass * 8;
      if (t < 256) sh_hist[t] = 0;
      __syncthreads();
      const unsigned long long prefix = limit_bits;
      const unsigned long long mask = (pass == 7) ? 0ull : (~0ull << (shift + 8));
#pragma unroll 1
      for (uint32_t k = t; k < total; k += kSelThreads) {
        const unsigned long long key = flat_key[k];
        if ((key & mask) == prefix) atomicAdd(&sh_hist[(key >> shift) & 255ull], 1u);
      }
      __syncthreads();
      if (w == 0) {   // parallel digit search: 8 bins per lane
        uint32_t c[8], s = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = sh_hist[lane * 8 + k]; s += c[k]; }
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const uint32_t excl = incl - s;
        const uint32_t rk = (uint32_t)sh_rank;
        __syncwarp();
        if (rk >= excl && rk < excl + s) {
          uint32_t run = excl;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (rk >= run && rk < run + c[k]) {
              sh_rank = (int)(rk - run);
              limit_bits = prefix | ((unsigned long long)(lane * 8 + k) << shift);
            }
            run += c[k];
          }
        }
      }
      __syncthreads();
    }
  }
  const double limit = __longlong_as_double((long long)limit_bits);
  // ---- members of the quantile bin with d2 <= limit (icp_fast.cc:497-498) -----------------
  double acc[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) acc[k] = 0.0;
#pragma unroll 1
  for (uint32_t k = t; k < total; k += kSelThreads) {
    const double d2 = __longlong_as_double((long long)flat_key[k]);
    if (d2 <= limit) {
      double px, py, pz; BucketPoint q; BucketNormal n;
      load_match(b, T, (int)flat_idx[k], px, py, pz, q, n);
      accumulate_match(acc, px, py, pz, q, n, d2);
    }
  }
  block_reduce_sums<kSelThreads>(acc, red, cand_sums);
  // ---- fixed-order reduction of the phase-B partials ---------------------------------------
  if (w < kNumSums) {
    double v = 0.0;
#pragma unroll 1
    for (int blk = lane; blk < nblocks_b; blk += 32) v += b.partials[(int64_t)blk * 32 + w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) solve.S[w] = v;
  }
#pragma unroll 1
  for (int k = t; k < 2 * kHistBins; k += kSelThreads) b.hist[k] = 0;   // hist and hist2
  __syncthreads();
  if (t < kNumSums) solve.S[t] += cand_sums[t];
  __syncthreads();
  if (w != 0) return;
  // ---- warp 0: solve + pose update -----------------------------------------------------------
  const double kept = solve.S[28];
  if (lane == 0) { st->limit = limit; st->kept = (long long)kept; }
  if (!(kept > 0.0)) {
    if (lane == 0) { st->status = -3; st->done = 1; }  // "no point to minimize", icp_fast.cc:114
    return;
  }
  const int path = solve_warp(solve, lane);
  if (lane == 0) update_pose(st, p, solve, T, path);
}

}  // namespace

void icp_finish_launch(const IcpBuffers& b, const IcpParams& p, int nblocks_b, cudaStream_t stream) {
  icp_finish_kernel<<<1, kSelThreads, 0, stream>>>(b, p, nblocks_b);
}

}  // namespace smb
