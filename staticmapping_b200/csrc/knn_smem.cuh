// libnabo-compatible 1-NN (KDTREE_LINEAR_HEAP, k = 1, allowSelfMatch, maxRadius = inf;
// call site registrators/icp_fast.cc:177-178) over the compact tree layout (KdCompact), with
// the node arrays resident in SHARED MEMORY.
//
// B200 formulation
//   * nodes are {cut f64}[h] + {dim u8}[h] in heap order.  Every CTA stages the TOP levels
//     (kKnnSmemLevels = 11: 2047 nodes, 18 KB — the part of the tree every query walks) into
//     shared memory with bulk async copies (cp.async.bulk.shared::cluster.global + mbarrier
//     complete_tx; SASS UBLKCP / SYNCS) while its threads fetch and transform their queries;
//     the deeper levels are read through the read-only path (L1).  Staging all 14 inner levels
//     of a ~107 k-point target (144 KB, one 1024-thread CTA per SM) was measured too: the same
//     57-60 us per launch, but 850 instead of ~1000 alignments/s with 16 alignments in flight,
//     because one CTA per SM leaves no room for the kernels of other alignments;
//   * leaves carry no payload: a leaf's bucket is (in-level index << (levels - level)) in the
//     padded bucket array, one bucket = x[8] y[8] z[8] = 12 aligned 16-byte loads;
//   * the heap index of a reached leaf encodes its whole path, so the far-side tests of a frame
//     are re-derived from it after the bucket, deepest level first, as the recursion would test
//     them (knn1_frames): no per-level bookkeeping during a descent, and the frame stack is
//     touched only by nested far visits.
// The visit ORDER and every comparison are those of libnabo's recurseKnn, so index sets and
// squared distances are bit-identical to the oracle for any epsilon (tests/test_gpu_icp.py).
#ifndef SM_B200_KNN_SMEM_CUH_
#define SM_B200_KNN_SMEM_CUH_

#include "common.cuh"
#include "kernels.h"

namespace smb {
namespace dev {

#ifndef SMB_KNN_THREADS
#define SMB_KNN_THREADS 256
#endif
#ifndef SMB_KNN_SMEM_LEVELS
#define SMB_KNN_SMEM_LEVELS 11
#endif
#ifndef SMB_KNN_LDG256
#define SMB_KNN_LDG256 1
#endif
#ifndef SMB_KNN_HALF_SCAN
#define SMB_KNN_HALF_SCAN 1
#endif
#ifndef SMB_KNN_Q_FROM_SMEM
#define SMB_KNN_Q_FROM_SMEM 0
#endif
#if SMB_KNN_Q_FROM_SMEM      // the query of a bucket scan re-read from its shared-memory column (frees six registers)
#define SMB_KNN_Q(t, reg, d) ((t).sq[(d) * kKnnCtaThreads])
#else
#define SMB_KNN_Q(t, reg, d) (reg)
#endif
#ifndef SMB_KNN_PAIR_PREFETCH
#define SMB_KNN_PAIR_PREFETCH 0
#endif
constexpr int kKnnCtaThreads = SMB_KNN_THREADS;        // threads (= traversal slots) per CTA
constexpr int kKnnSmemLevels = SMB_KNN_SMEM_LEVELS;    // tree levels staged in shared memory: 2^11 * 9 B = 18 KB
constexpr int kKnnCtasPerSm = 1024 / kKnnCtaThreads;   // 64 registers per thread -> 1024 threads per SM

struct SmemTree {
  const double* s_cut;      // shared memory
  const uint8_t* s_dim;     // shared memory
  int n_smem;               // heap indices < n_smem are resident in shared memory
  const double* g_cut;      // global (all nodes)
  const uint8_t* g_dim;
  const double2* g_node;    // global, packed {cut, dim} of heap node h at slot h + 1 (one 16-byte load per node)
  const double* pb;         // padded buckets
  int levels;
  double* sq;               // this thread's query: sq[d * kKnnCtaThreads] (shared memory, d = 0..2)
  double* so;               // this thread's off[] of the recursion, same layout
};

// ---- mbarrier + bulk async copy (TMA engine, non-tensor form) --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}

// shared-memory footprint of the staged tree: cut[slots] + dim[slots] + barrier + 16 doubles
__host__ __device__ __forceinline__ int knn_smem_slots(int levels) {
  const int ls = levels < kKnnSmemLevels ? levels : kKnnSmemLevels;
  const int n = 1 << ls;
  return n < 16 ? 16 : n;
}
__host__ __device__ __forceinline__ size_t knn_smem_bytes(int levels) {
  // nodes, barrier (+pad), 16 doubles, per-thread query / offset columns
  return (size_t)knn_smem_slots(levels) * 9 + 16 + 16 * sizeof(double) + 6 * (size_t)kKnnCtaThreads * sizeof(double);
}

// Called by ALL threads of the CTA (contains __syncthreads).  Thread 0 arms the barrier and
// issues the copies; everyone may then do independent work and must call mbar_wait(bar, 0)
// before the first tree access.
__device__ __forceinline__ SmemTree stage_tree(const KdCompact& t, unsigned char* smem, uint64_t** bar_out,
                                               double** extra_out) {
  const int slots = knn_smem_slots(t.levels);
  double* s_cut = reinterpret_cast<double*>(smem);
  uint8_t* s_dim = smem + (size_t)slots * 8;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)slots * 9);
  *extra_out = reinterpret_cast<double*>(bar + 2);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_proxy_async_smem(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t cut_bytes = (uint32_t)slots * 8u, dim_bytes = (uint32_t)slots;
    mbar_arrive_expect_tx(bar, cut_bytes + dim_bytes);
    constexpr uint32_t kChunk = 32768u;
    for (uint32_t off = 0; off < cut_bytes; off += kChunk)
      bulk_copy_g2s(smem + off, reinterpret_cast<const unsigned char*>(t.cut) + off,
                    cut_bytes - off < kChunk ? cut_bytes - off : kChunk, bar);
    bulk_copy_g2s(s_dim, t.dim, dim_bytes, bar);
  }
  *bar_out = bar;
  SmemTree st;
  st.s_cut = s_cut; st.s_dim = s_dim; st.n_smem = slots;
  st.g_cut = t.cut; st.g_dim = t.dim; st.g_node = t.node; st.pb = t.pb; st.levels = t.levels;
  st.sq = *extra_out + 16 + threadIdx.x;
  st.so = st.sq + 3 * kKnnCtaThreads;
  return st;
}

template <bool kAllSmem>
__device__ __forceinline__ void tree_node(const SmemTree& t, int h, double& cut, int& dim) {
  if (kAllSmem || h < t.n_smem) { cut = t.s_cut[h]; dim = t.s_dim[h]; }
  else { cut = __ldg(t.g_cut + h); dim = __ldg(t.g_dim + h); }
}

__device__ __forceinline__ double2 ldg_nc_f64x2(const void* p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// A node of the levels below the staged ones: cut and dimension in ONE 16-byte load (two dependent
// loads of two arrays before).
__device__ __forceinline__ void ldg_node(const double2* __restrict__ g_node, int h, double& cut, int& dim) {
  const double2 v = ldg_nc_f64x2(g_node + h + 1);
  cut = v.x;
  dim = (int)__double_as_longlong(v.y);
}

struct Double4 { double a, b, c, d; };
__device__ __forceinline__ Double4 ldg_nc_f64x4(const void* p) {      // LDG.E.256 (sm_100), 32-byte aligned
  Double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v.a), "=d"(v.b), "=d"(v.c), "=d"(v.d) : "l"(p));
  return v;
}

// One padded bucket: 6 32-byte loads, 8 squared distances in the reference's operation order, then
// a tournament on the bit patterns in which the lower index wins ties (libnabo walks the bucket in
// order and replaces the head on a strict '<').  Padding entries are +inf: their distance is +inf
// or NaN, never below the head.  SMB_KNN_HALF_SCAN: four points at a time (half the registers).
__device__ __forceinline__ unsigned long long dist_key(double qx, double qy, double qz, double x, double y, double z) {
  const double dx = dsub(qx, x), dy = dsub(qy, y), dz = dsub(qz, z);
  return (unsigned long long)__double_as_longlong(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)));
}

__device__ __forceinline__ void scan_bucket(const double* __restrict__ pb, int bucket, double qx, double qy,
                                            double qz, double& head, int& best) {
  const char* base = reinterpret_cast<const char*>(pb + (int64_t)bucket * 24);
#if SMB_KNN_HALF_SCAN
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const Double4 vx = ldg_nc_f64x4(base + 32 * half), vy = ldg_nc_f64x4(base + 64 + 32 * half),
                  vz = ldg_nc_f64x4(base + 128 + 32 * half);
    unsigned long long k0 = dist_key(qx, qy, qz, vx.a, vy.a, vz.a), k1 = dist_key(qx, qy, qz, vx.b, vy.b, vz.b);
    unsigned long long k2 = dist_key(qx, qy, qz, vx.c, vy.c, vz.c), k3 = dist_key(qx, qy, qz, vx.d, vy.d, vz.d);
    int a0 = 0, a2 = 2;
    if (k1 < k0) { k0 = k1; a0 = 1; }
    if (k3 < k2) { k2 = k3; a2 = 3; }
    if (k2 < k0) { k0 = k2; a0 = a2; }
    if (k0 < (unsigned long long)__double_as_longlong(head)) {
      head = __longlong_as_double((long long)k0);
      best = bucket * 8 + 4 * half + a0;
    }
  }
#else
  unsigned long long key[8];
#if SMB_KNN_LDG256
  Double4 v[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = ldg_nc_f64x4(base + 32 * k);
#else
  double2 v[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) v[k] = ldg_nc_f64x2(base + 16 * k);
#endif
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#if SMB_KNN_LDG256
    const Double4 &vx = v[k >> 2], &vy = v[2 + (k >> 2)], &vz = v[4 + (k >> 2)];
    const double x = (k & 3) == 0 ? vx.a : (k & 3) == 1 ? vx.b : (k & 3) == 2 ? vx.c : vx.d;
    const double y = (k & 3) == 0 ? vy.a : (k & 3) == 1 ? vy.b : (k & 3) == 2 ? vy.c : vy.d;
    const double z = (k & 3) == 0 ? vz.a : (k & 3) == 1 ? vz.b : (k & 3) == 2 ? vz.c : vz.d;
#else
    const double x = (k & 1) ? v[k >> 1].y : v[k >> 1].x;
    const double y = (k & 1) ? v[4 + (k >> 1)].y : v[4 + (k >> 1)].x;
    const double z = (k & 1) ? v[8 + (k >> 1)].y : v[8 + (k >> 1)].x;
#endif
    key[k] = dist_key(qx, qy, qz, x, y, z);
  }
  int arg[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) arg[k] = k;
#pragma unroll
  for (int step = 1; step < 8; step <<= 1) {
#pragma unroll
    for (int k = 0; k < 8; k += 2 * step) {
      const bool take = key[k + step] < key[k];      // strict: the lower index keeps a tie
      key[k] = take ? key[k + step] : key[k];
      arg[k] = take ? arg[k + step] : arg[k];
    }
  }
  if (key[0] < (unsigned long long)__double_as_longlong(head)) {
    head = __longlong_as_double((long long)key[0]);
    best = bucket * 8 + arg[0];
  }
#endif
}

constexpr int kKnnMaxStack = 32;
constexpr int kKnnMaxLevels = 24;        // sm_api rejects deeper trees

// The coordinate of the query / the recursion's off[] along a node's cut dimension are read from
// per-thread shared-memory columns indexed by that dimension (one LDS instead of a three-way select
// on register pairs: the descent loop of a far visit shrank from 69 to ~30 SASS instructions a level).
struct KnnStackEntry {
  double rd;
  double o[3];
  int h, pad;
};

// recurseKnn on the subtree rooted at heap node h with the recursion's rd at entry; its off[] is in
// t.so.  Near child first; a far child is pushed if it passes rd_new*(1+eps)^2 < head now (the head
// only shrinks) and re-tested when popped, which is when the recursion tests it.  Far subtrees start
// deep in the tree: nodes come through the read-only path (L1).
__device__ __forceinline__ void visit_subtree(const SmemTree& t, double qx, double qy, double qz, double me2,
                                              int h, double rd, double& head, int& best) {
  KnnStackEntry stack[kKnnMaxStack];
  int sp = 0;
  while (true) {
    int l = 31 - __clz(h + 1);
#if SMB_KNN_PAIR_PREFETCH
    double cut = 0.0; int cd = 3;
    if (l < t.levels) ldg_node(t.g_node, h, cut, cd);
#endif
    while (l < t.levels) {
#if SMB_KNN_PAIR_PREFETCH
      if (cd == 3) break;
      Double4 ch = {0.0, 0.0, 0.0, 0.0};                       // both children, fetched while this node is decided
      if (l + 1 < t.levels) ch = ldg_nc_f64x4(t.g_node + 2 * (h + 1));
#else
      double cut; int cd;
      ldg_node(t.g_node, h, cut, cd);
      if (cd == 3) break;
#endif
      const double q = t.sq[cd * kKnnCtaThreads];
      const double old_off = t.so[cd * kKnnCtaThreads];
      const double new_off = dsub(q, cut);
      const double rd_new = dadd(rd, dadd(-dmul(old_off, old_off), dmul(new_off, new_off)));
      const int right = q > cut ? 1 : 0;              // == (q - cut > 0) for IEEE doubles
      if (dmul(rd_new, me2) < head && sp < kKnnMaxStack) {
        KnnStackEntry& e = stack[sp++];
        e.rd = rd_new;
        e.o[0] = t.so[0]; e.o[1] = t.so[kKnnCtaThreads]; e.o[2] = t.so[2 * kKnnCtaThreads];
        e.o[cd] = new_off;
        e.h = 2 * h + 2 - right;
      }
      h = 2 * h + 1 + right;
      ++l;
#if SMB_KNN_PAIR_PREFETCH
      cut = right ? ch.c : ch.a;
      cd = (int)__double_as_longlong(right ? ch.d : ch.b);
#endif
    }
    scan_bucket(t.pb, (h + 1 - (1 << l)) << (t.levels - l), SMB_KNN_Q(t, qx, 0), SMB_KNN_Q(t, qy, 1), SMB_KNN_Q(t, qz, 2), head, best);
    bool found = false;
    while (sp > 0) {
      const KnnStackEntry& e = stack[--sp];
      if (dmul(e.rd, me2) < head) {
        h = e.h; rd = e.rd;
        t.so[0] = e.o[0]; t.so[kKnnCtaThreads] = e.o[1]; t.so[2 * kKnnCtaThreads] = e.o[2];
        found = true;
        break;
      }
    }
    if (!found) break;
  }
}

// First visit of a query: the descent from the root (staged levels from shared memory, the rest one
// packed record per level), the first bucket, and the mask of the path levels whose far side passes
// with the head that bucket left.  Path node of level a: ((hp1) >> (ll-a)) - 1.  rd = 0 and off = 0
// along the whole root path, so rd_new(a) = 0 + (-(0*0) + new_off^2) = new_off^2 exactly.
template <bool kAllSmem>
__device__ __forceinline__ void knn_root_visit(const SmemTree& t, double qx, double qy, double qz, double me2,
                                               double& head, int& best, int& hp1_out, int& ll_out, uint32_t& mask_out) {
  head = __longlong_as_double(0x7ff0000000000000ll);
  best = -1;
  t.sq[0] = qx; t.sq[kKnnCtaThreads] = qy; t.sq[2 * kKnnCtaThreads] = qz;
  const int lsm = kAllSmem ? t.levels : min(t.levels, kKnnSmemLevels);    // levels staged in shared memory
  // lb[a]: the squared distance to the cutting plane of path level a, rounded DOWN to float while the
  // descent has the node in hand.  RN(lb * me2) <= RN(off^2 * me2), so "lb * me2 < head" holds whenever the
  // exact test does: the mask below is a superset of the levels that pass, and every level of it is
  // re-tested exactly (node re-read) at the moment the recursion would test it.  The array lives in local
  // memory; all lanes of a warp store / load the same level together (one 128-byte line per access).
  float lb[kKnnMaxLevels];
  int h = 0, l = 0;
  bool leaf = false;
  while (l < lsm) {
    const int cd = t.s_dim[h];
    if (cd == 3) { leaf = true; break; }
    const double q = t.sq[cd * kKnnCtaThreads], off = dsub(q, t.s_cut[h]);
    lb[l] = __double2float_rd(dmul(off, off));
    h = 2 * h + 1 + (off > 0.0 ? 1 : 0);              // == (q > cut) for IEEE doubles
    ++l;
  }
  if (!kAllSmem && !leaf) {
    while (l < t.levels) {
      double cut; int cd;
      ldg_node(t.g_node, h, cut, cd);
      if (cd == 3) break;
      const double q = t.sq[cd * kKnnCtaThreads], off = dsub(q, cut);
      lb[l] = __double2float_rd(dmul(off, off));
      h = 2 * h + 1 + (off > 0.0 ? 1 : 0);
      ++l;
    }
  }
  scan_bucket(t.pb, (h + 1 - (1 << l)) << (t.levels - l), qx, qy, qz, head, best);
  const int hp1 = h + 1, ll = l;
  uint32_t mask = 0u;
#pragma unroll 4
  for (int a = 0; a < ll; ++a)
    if (dmul((double)lb[a], me2) < head) mask |= 1u << a;
  hp1_out = hp1; ll_out = ll; mask_out = mask;
}

// best: padded entry index (bucket * 8 + k) or -1; d2: squared distance (+inf if none)
template <bool kAllSmem>
__device__ __forceinline__ void knn1_smem(const SmemTree& t, double qx, double qy, double qz, double me2,
                                          int& best_out, double& d2_out) {
  double head; int best, hp1, ll; uint32_t mask;
  knn_root_visit<kAllSmem>(t, qx, qy, qz, me2, head, best, hp1, ll, mask);
  // the levels of the mask deepest first, each re-tested with the head of that moment
  while (mask != 0u) {
    const int a = 31 - __clz(mask);
    mask &= ~(1u << a);
    const int n = (hp1 >> (ll - a)) - 1;
    double cut; int cd;
    ldg_node(t.g_node, n, cut, cd);
    const double off = dsub(t.sq[cd * kKnnCtaThreads], cut);
    const double rd_new = dmul(off, off);
    if (dmul(rd_new, me2) < head) {
      t.so[0] = 0.0; t.so[kKnnCtaThreads] = 0.0; t.so[2 * kKnnCtaThreads] = 0.0;
      t.so[cd * kKnnCtaThreads] = off;
      const int near_p1 = hp1 >> (ll - a - 1);       // path node of level a+1, plus one
      visit_subtree(t, qx, qy, qz, me2, (near_p1 ^ 1) - 1, rd_new, head, best);
    }
  }
  best_out = best;
  d2_out = head;
}

// ---- far visits of a batch of queries, one WARP working through a list ---------------------------
// A query whose root visit left far-side candidates (knn_root_visit's mask != 0) is parked as a
// KnnItem; the lanes of the warp then pull items from the list.  Every pass of the loop below is ONE
// bucket visit per busy lane — the next subtree of the lane's query (a nested far child popped from its
// stack, else the next root-path level of the mask, both re-tested with the head of that moment exactly
// as recurseKnn tests them), the descent into it with far children pushed, the bucket — and a lane
// whose query is finished takes the next item in the same pass.  The per-query visit order is the
// recursion's; only WHICH lane runs a query and when is different, so results are bit-identical.
// Static one-query-per-lane scheduling leaves 13 of 32 lanes busy (the visit count per query is 1..29,
// mean 2.9); here the bucket scans and descents of a pass run with most lanes busy.
struct KnnItem { int i, hp1, ll; uint32_t mask; };      // 16 bytes

template <typename LoadQuery, typename StoreResult>
__device__ __forceinline__ void knn_far_phase(const SmemTree& t, double me2, const int4* __restrict__ items,
                                              int n_items, LoadQuery load_query, StoreResult store_result) {
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  KnnStackEntry stack[kKnnMaxStack];
  int sp = 0, next = 0;
  bool have = false;
  int qi = 0, hp1 = 1, ll = 0, best = -1;
  uint32_t mask = 0u;
  double head = 0.0, qx = 0.0, qy = 0.0, qz = 0.0;
  while (true) {
    const unsigned need = __ballot_sync(0xffffffffu, !have);
    if (!have) {
      const int idx = next + __popc(need & lt);
      if (idx < n_items) {
        const int4 it = __ldcg(items + idx);
        qi = it.x; hp1 = it.y; ll = it.z; mask = (uint32_t)it.w;
        load_query(qi, qx, qy, qz, head, best);
        t.sq[0] = qx; t.sq[kKnnCtaThreads] = qy; t.sq[2 * kKnnCtaThreads] = qz;
        sp = 0;
        have = true;
      }
    }
    next += __popc(need);
    if (!__any_sync(0xffffffffu, have)) break;
    bool go = false;
    int h = 0;
    double rd = 0.0;
    if (have) {
      while (sp > 0) {                                 // nested far children first (the recursion is inside them)
        const KnnStackEntry& e = stack[--sp];
        if (dmul(e.rd, me2) < head) {
          h = e.h; rd = e.rd;
          t.so[0] = e.o[0]; t.so[kKnnCtaThreads] = e.o[1]; t.so[2 * kKnnCtaThreads] = e.o[2];
          go = true;
          break;
        }
      }
      while (!go && mask != 0u) {                      // then the root path, deepest level first
        const int a = 31 - __clz(mask);
        mask &= ~(1u << a);
        const int n = (hp1 >> (ll - a)) - 1;
        double cut; int cd;
        ldg_node(t.g_node, n, cut, cd);
        const double off = dsub(t.sq[cd * kKnnCtaThreads], cut);
        const double rd_new = dmul(off, off);
        if (dmul(rd_new, me2) < head) {
          t.so[0] = 0.0; t.so[kKnnCtaThreads] = 0.0; t.so[2 * kKnnCtaThreads] = 0.0;
          t.so[cd * kKnnCtaThreads] = off;
          h = ((hp1 >> (ll - a - 1)) ^ 1) - 1;         // sibling of the path node of level a + 1
          rd = rd_new;
          go = true;
        }
      }
      if (!go) { store_result(qi, best, head); have = false; }
    }
    if (go) {
      int l = 31 - __clz(h + 1);
      while (l < t.levels) {
        double cut; int cd;
        ldg_node(t.g_node, h, cut, cd);
        if (cd == 3) break;
        const double q = t.sq[cd * kKnnCtaThreads];
        const double old_off = t.so[cd * kKnnCtaThreads];
        const double new_off = dsub(q, cut);
        const double rd_new = dadd(rd, dadd(-dmul(old_off, old_off), dmul(new_off, new_off)));
        const int right = q > cut ? 1 : 0;
        if (dmul(rd_new, me2) < head && sp < kKnnMaxStack) {
          KnnStackEntry& e = stack[sp++];
          e.rd = rd_new;
          e.o[0] = t.so[0]; e.o[1] = t.so[kKnnCtaThreads]; e.o[2] = t.so[2 * kKnnCtaThreads];
          e.o[cd] = new_off;
          e.h = 2 * h + 2 - right;
        }
        h = 2 * h + 1 + right;
        ++l;
      }
      scan_bucket(t.pb, (h + 1 - (1 << l)) << (t.levels - l), qx, qy, qz, head, best);
    }
  }
}

// One CTA's range [begin, end) in batches of 256 * qpt queries (per_cta is a multiple of the CTA size):
//   (1) every thread loads its queries (one per pass, consecutive threads = consecutive queries) and does
//       their root visits in lockstep with its warp.  A query without far-side candidates is final here;
//   (2) the others are parked (park: head and best of the moment; a 16-byte item in the warp's slice of
//       items_all, which lies inside the batch's own index range) and the warp works through its list
//       (knn_far_phase).
// load_q(i, x, y, z), park(i, best, head), unpark(i, head&, best&), finish(i, best, head).
constexpr int kKnnMaxQpt = 4;
template <bool kAllSmem, typename LoadQ, typename Park, typename Unpark, typename Finish>
__device__ __forceinline__ void knn_batch_cta(const SmemTree& tree, int begin, int end, int per_cta, double me2,
                                              int4* __restrict__ items_all, LoadQ load_q, Park park,
                                              Unpark unpark, Finish finish) {
  const int warp = threadIdx.x >> 5;
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  const int qpt = min(kKnnMaxQpt, per_cta / kKnnCtaThreads);
  for (int base = begin; base < end; base += kKnnCtaThreads * qpt) {
    const int qb = min(qpt, (begin + per_cta - base) / kKnnCtaThreads);
    int4* items = items_all + base + warp * 32 * qb;
    int count = 0;
    for (int j = 0; j < qb; ++j) {
      const int i = base + j * kKnnCtaThreads + threadIdx.x;
      bool far = false;
      int4 item = make_int4(0, 0, 0, 0);
      if (i < end) {
        double px, py, pz, head;
        int best, hp1, ll;
        uint32_t mask;
        load_q(i, px, py, pz);
        knn_root_visit<kAllSmem>(tree, px, py, pz, me2, head, best, hp1, ll, mask);
        if (mask == 0u) {
          finish(i, best, head);
        } else {
          park(i, best, head);
          far = true;
          item = make_int4(i, hp1, ll, (int)mask);
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, far);
      if (far) items[count + __popc(m & lt)] = item;
      count += __popc(m);
    }
    __syncwarp();
    knn_far_phase(tree, me2, items, count,
                  [&](int i, double& qx, double& qy, double& qz, double& head, int& best) {
                    load_q(i, qx, qy, qz);
                    unpark(i, head, best);
                  },
                  finish);
    __syncwarp();
  }
}

}  // namespace dev
}  // namespace smb

#endif  // SM_B200_KNN_SMEM_CUH_
