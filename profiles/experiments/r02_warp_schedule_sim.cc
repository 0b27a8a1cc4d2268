// warp-level cost simulation of scheduling policies for the libnabo traversal
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <numeric>
#include <cstdint>
using namespace std;
struct Node { double cut; int dim; int first, count; };
int N; vector<double> P; vector<Node> nodes; vector<int> idx;
void build(int h, int first, int count, double mn[3], double mx[3]) {
  if ((int)nodes.size() <= h) nodes.resize(h + 1, Node{0, -1, 0, 0});
  if (count <= 8) { nodes[h] = Node{0, 3, first, count}; return; }
  int dim = 0; double mv = 0; for (int d = 0; d < 3; ++d) if (mx[d] - mn[d] > mv) { mv = mx[d] - mn[d]; dim = d; }
  int right = count / 2, left = count - right;
  nth_element(idx.begin() + first, idx.begin() + first + left, idx.begin() + first + count,
              [&](int a, int b) { return P[3 * a + dim] < P[3 * b + dim] || (P[3 * a + dim] == P[3 * b + dim] && a < b); });
  double cut = P[3 * idx[first + left] + dim];
  nodes[h] = Node{cut, dim, first, count};
  double lmx[3] = {mx[0], mx[1], mx[2]}, rmn[3] = {mn[0], mn[1], mn[2]};
  lmx[dim] = cut; rmn[dim] = cut;
  build(2 * h + 1, first, left, mn, lmx); build(2 * h + 2, first + left, right, rmn, mx);
}
double ME2 = 4.16 * 4.16;
struct Visit { int8_t desc, masklv; };   // descent levels, levels of the new frame to test
vector<vector<Visit>> V;
void scan(int h, const double* q, double& head) {
  const Node& n = nodes[h];
  for (int i = 0; i < n.count; ++i) { const double* p = &P[3 * idx[n.first + i]]; double d = 0; for (int r = 0; r < 3; ++r) { double t = q[r] - p[r]; d += t * t; } if (d < head) head = d; }
}
// visit subtree from node h at level l: descend near path to leaf, scan, then far candidates deepest first
void visit(int qi, int h, int l, const double* q, double rd, double off[3], double& head) {
  int path[32]; int lev = l; int hh = h; int n0 = 0;
  while (nodes[hh].dim != 3) { path[n0++] = hh; double no = q[nodes[hh].dim] - nodes[hh].cut; hh = no > 0 ? 2 * hh + 2 : 2 * hh + 1; ++lev; }
  scan(hh, q, head);
  V[qi].push_back(Visit{(int8_t)n0, (int8_t)n0});
  for (int k = n0 - 1; k >= 0; --k) {
    const Node& n = nodes[path[k]]; int cd = n.dim; double old = off[cd], no = q[cd] - n.cut;
    double rdn = rd + (-old * old + no * no);
    if (rdn * ME2 < head) { int far = no > 0 ? 2 * path[k] + 1 : 2 * path[k] + 2; off[cd] = no; visit(qi, far, l + k + 1, q, rdn, off, head); off[cd] = old; }
  }
}
// cost model (thread instructions per warp-iteration): unified visit
double A_FIX = 120, A_DESC = 13, A_BUCKET = 150, A_MASK = 30;
double itercost(int maxdesc, int maxmask) { return A_FIX + A_DESC * maxdesc + A_BUCKET + A_MASK * maxmask; }
int main(int argc, char** argv) {
  FILE* f = fopen("/tmp/ana/tgt.bin", "rb"); fseek(f, 0, SEEK_END); N = ftell(f) / 24; fseek(f, 0, SEEK_SET); P.resize(3 * N); if (fread(P.data(), 8, 3 * N, f)) {} fclose(f);
  f = fopen(argc > 1 ? argv[1] : "/tmp/ana/src.bin", "rb"); fseek(f, 0, SEEK_END); int M = ftell(f) / 24; fseek(f, 0, SEEK_SET); vector<double> Q(3 * M); if (fread(Q.data(), 8, 3 * M, f)) {} fclose(f);
  idx.resize(N); iota(idx.begin(), idx.end(), 0);
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < N; ++i) for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], P[3 * i + d]); mx[d] = max(mx[d], P[3 * i + d]); }
  build(0, 0, N, mn, mx);
  // Morton order of queries (0.25 m lattice like the kernel)
  vector<pair<uint64_t,int>> mo(M);
  auto spread = [](uint32_t v) { uint64_t x = v & 1023; x = (x | (x << 16)) & 0x030000ff; x = (x | (x << 8)) & 0x0300f00f; x = (x | (x << 4)) & 0x030c30c3; x = (x | (x << 2)) & 0x09249249; return x; };
  for (int i = 0; i < M; ++i) { uint32_t ix = (int)floor(Q[3*i]*4), iy = (int)floor(Q[3*i+1]*4), iz = (int)floor(Q[3*i+2]*4); mo[i] = {spread(ix) | (spread(iy) << 1) | (spread(iz) << 2), i}; }
  sort(mo.begin(), mo.end());
  V.resize(M);
  for (int k = 0; k < M; ++k) { int i = mo[k].second; double off[3] = {0,0,0}; double head = INFINITY; visit(k, 0, 0, &Q[3*i], 0, off, head); }
  double tot_visits = 0; for (auto& v : V) tot_visits += v.size();
  printf("queries %d visits %.0f (%.2f per query)\n", M, tot_visits, tot_visits / M);
  // ideal: every visit at its own cost, perfectly packed
  double ideal = 0; for (auto& v : V) for (auto& x : v) ideal += itercost(x.desc, x.masklv);
  printf("ideal (no divergence) warp-instr: %.2f M\n", ideal / 32 / 1e6);
  // policy 1: static, one query per lane, warp iterates until all done
  { double c = 0; long iters = 0;
    for (int w = 0; w < M; w += 32) { size_t mxv = 0; for (int l = w; l < min(M, w + 32); ++l) mxv = max(mxv, V[l].size());
      for (size_t r = 0; r < mxv; ++r) { int md = 0, mm = 0; for (int l = w; l < min(M, w + 32); ++l) if (r < V[l].size()) { md = max(md, (int)V[l][r].desc); mm = max(mm, (int)V[l][r].masklv); } c += itercost(md, mm); ++iters; } }
    printf("static per-thread: %.2f M warp-instr, %ld warp-iterations\n", c / 1e6, iters); }
  // policy 2: refill when idle lanes >= TH (warp owns a contiguous chunk of CH queries)
  for (int CH : {128, 256, 1024}) for (int TH : {1, 8, 16, 24}) {
    double c = 0; long iters = 0; long maxchain_iters = 0;
    for (int w0 = 0; w0 < M; w0 += CH) {
      int next = w0, endq = min(M, w0 + CH); int lane_q[32], lane_r[32]; for (int l = 0; l < 32; ++l) lane_q[l] = -1;
      long it_here = 0;
      while (true) {
        int idle = 0; for (int l = 0; l < 32; ++l) if (lane_q[l] < 0) ++idle;
        if (next < endq && (idle >= TH || idle == 32)) { for (int l = 0; l < 32 && next < endq; ++l) if (lane_q[l] < 0) { lane_q[l] = next++; lane_r[l] = 0; } }
        int md = 0, mm = 0, act = 0;
        for (int l = 0; l < 32; ++l) if (lane_q[l] >= 0) { const Visit& x = V[lane_q[l]][lane_r[l]]; md = max(md, (int)x.desc); mm = max(mm, (int)x.masklv); ++act; }
        if (!act) break;
        c += itercost(md, mm) + 15; ++iters; ++it_here;
        for (int l = 0; l < 32; ++l) if (lane_q[l] >= 0) { if (++lane_r[l] >= (int)V[lane_q[l]].size()) lane_q[l] = -1; }
      }
      maxchain_iters = max(maxchain_iters, it_here);
    }
    printf("refill CH=%4d TH=%2d: %.2f M warp-instr, %ld iterations, longest warp %ld iters\n", CH, TH, c / 1e6, iters, maxchain_iters);
  }
  // policy 3: warp-local rounds with pool (compaction each pass), chunk CH, +60 instr state traffic per pass
  for (int CH : {128, 256}) {
    double c = 0; long iters = 0;
    for (int w0 = 0; w0 < M; w0 += CH) {
      vector<pair<int,int>> live; for (int q = w0; q < min(M, w0 + CH); ++q) live.push_back({q, 0});
      while (!live.empty()) { vector<pair<int,int>> nxt;
        for (size_t s = 0; s < live.size(); s += 32) { int md = 0, mm = 0; for (size_t k = s; k < min(live.size(), s + 32); ++k) { const Visit& x = V[live[k].first][live[k].second]; md = max(md, (int)x.desc); mm = max(mm, (int)x.masklv); }
          c += itercost(md, mm) + 60; ++iters; }
        for (auto& p : live) if (p.second + 1 < (int)V[p.first].size()) nxt.push_back({p.first, p.second + 1});
        live.swap(nxt); }
    }
    printf("warp-local rounds CH=%d: %.2f M warp-instr, %ld passes\n", CH, c / 1e6, iters);
  }
  return 0;
}
