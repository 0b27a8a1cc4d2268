// offline statistics of the libnabo traversal on the config-2 workload
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <numeric>
using namespace std;
struct Node { double cut; int dim; int first, count; };
int N; vector<double> P; vector<Node> nodes; vector<int> order; int LEV;
void build(int h, int first, int count, double mn[3], double mx[3], vector<int>& idx) {
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
  build(2 * h + 1, first, left, mn, lmx, idx);
  build(2 * h + 2, first + left, right, rmn, mx, idx);
}
struct Stats { int rounds = 0, c0 = 0, vroot = 0, nested = 0, maxdepth = 0, waves = 0; };
double ME2 = 4.16 * 4.16;
vector<int> idx;
void scan(int h, const double* q, double& head) {
  const Node& n = nodes[h];
  for (int i = 0; i < n.count; ++i) { const double* p = &P[3 * idx[n.first + i]]; double d = 0; for (int r = 0; r < 3; ++r) { double t = q[r] - p[r]; d += t * t; } if (d < head) head = d; }
}
// returns number of waves needed for this subtree under the speculative scheme
void rec(int h, const double* q, double rd, double off[3], double& head, Stats& s, int depth, bool rootframe) {
  const Node& n = nodes[h];
  if (n.dim == 3) { scan(h, q, head); s.rounds++; return; }
  int cd = n.dim; double old = off[cd], no = q[cd] - n.cut;
  int near = no > 0 ? 2 * h + 2 : 2 * h + 1, far = no > 0 ? 2 * h + 1 : 2 * h + 2;
  rec(near, q, rd, off, head, s, depth, rootframe);
  rd += -old * old + no * no;
  if (rd * ME2 < head) {
    if (rootframe) s.vroot++; else s.nested++;
    s.maxdepth = max(s.maxdepth, depth + 1);
    off[cd] = no; rec(far, q, rd, off, head, s, depth + 1, false); off[cd] = old;
  }
}
int main(int argc, char** argv) {
  FILE* f = fopen("/tmp/ana/tgt.bin", "rb"); fseek(f, 0, SEEK_END); N = ftell(f) / 24; fseek(f, 0, SEEK_SET); P.resize(3 * N); fread(P.data(), 8, 3 * N, f); fclose(f);
  f = fopen(argc > 1 ? argv[1] : "/tmp/ana/src.bin", "rb"); fseek(f, 0, SEEK_END); int M = ftell(f) / 24; fseek(f, 0, SEEK_SET); vector<double> Q(3 * M); fread(Q.data(), 8, 3 * M, f); fclose(f);
  idx.resize(N); iota(idx.begin(), idx.end(), 0);
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = 0; i < N; ++i) for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], P[3 * i + d]); mx[d] = max(mx[d], P[3 * i + d]); }
  build(0, 0, N, mn, mx, idx);
  vector<Stats> st(M);
  vector<double> heads(M);
  for (int i = 0; i < M; ++i) {
    const double* q = &Q[3 * i]; double off[3] = {0, 0, 0}; double head = INFINITY;
    // c0: root candidates with head0
    { int h = 0; while (nodes[h].dim != 3) { double no = q[nodes[h].dim] - nodes[h].cut; h = no > 0 ? 2 * h + 2 : 2 * h + 1; }
      double h0 = INFINITY; scan(h, q, h0);
      int hh = 0; while (nodes[hh].dim != 3) { double no = q[nodes[hh].dim] - nodes[hh].cut; if (no * no * ME2 < h0) st[i].c0++; hh = no > 0 ? 2 * hh + 2 : 2 * hh + 1; } }
    rec(0, q, 0, off, head, st[i], 0, true);
    heads[i] = head;
  }
  auto pct = [&](auto get, const char* name) {
    vector<int> v(M); for (int i = 0; i < M; ++i) v[i] = get(st[i]); sort(v.begin(), v.end());
    double mean = accumulate(v.begin(), v.end(), 0.0) / M;
    printf("%-10s mean %.2f p50 %d p90 %d p99 %d p999 %d max %d\n", name, mean, v[M / 2], v[M * 9 / 10], v[M * 99 / 100], v[(long)M * 999 / 1000], v[M - 1]);
  };
  pct([](const Stats& s) { return s.rounds; }, "rounds");
  pct([](const Stats& s) { return s.c0; }, "c0");
  pct([](const Stats& s) { return s.vroot; }, "vroot");
  pct([](const Stats& s) { return s.nested; }, "nested");
  pct([](const Stats& s) { return s.maxdepth; }, "maxdepth");
  // rounds vs final distance
  vector<int> ord(M); iota(ord.begin(), ord.end(), 0); sort(ord.begin(), ord.end(), [&](int a, int b) { return heads[a] < heads[b]; });
  for (int dec = 0; dec < 10; ++dec) { double s = 0, c = 0; int mx2 = 0; for (int k = dec * M / 10; k < (dec + 1) * M / 10; ++k) { s += st[ord[k]].rounds; c += st[ord[k]].c0; mx2 = max(mx2, st[ord[k]].rounds); }
    printf("decile %d: d=%.3f mean rounds %.2f mean c0 %.2f max rounds %d\n", dec, sqrt(heads[ord[(dec + 1) * M / 10 - 1]]), s / (M / 10), c / (M / 10), mx2); }
  // warp-level: consecutive 32 (not Morton sorted here)
  return 0;
}
