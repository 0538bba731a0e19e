#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
struct Edge12 { float w; int a, b; };
struct UfElt { int rank, p, size; float thr; };
static inline int uf_find(UfElt* u, int x) { int y = x; while (y != u[y].p) y = u[y].p; u[x].p = y; return y; }
static inline void uf_join(UfElt* u, int x, int y) {
  if (u[x].rank > u[y].rank) { u[y].p = x; u[x].size += u[y].size; }
  else { u[x].p = y; u[y].size += u[x].size; if (u[x].rank == u[y].rank) u[y].rank++; }
}
template <int PF, int PF2>
void kruskal(const Edge12* e, size_t nE, size_t nV, float c, std::vector<UfElt>& u) {
  u.resize(nV);
  for (size_t i = 0; i < nV; ++i) { u[i].rank = 0; u[i].size = 1; u[i].p = (int)i; u[i].thr = c; }
  UfElt* U = u.data();
  for (size_t i = 0; i < nE; ++i) {
    if (i + PF < nE) { __builtin_prefetch(U + e[i + PF].a); __builtin_prefetch(U + e[i + PF].b); }
    if (PF2 && i + PF2 < nE) { __builtin_prefetch(U + U[e[i + PF2].a].p); __builtin_prefetch(U + U[e[i + PF2].b].p); }
    int a = uf_find(U, e[i].a), b = uf_find(U, e[i].b);
    if (a != b && e[i].w <= U[a].thr && e[i].w <= U[b].thr) {
      uf_join(U, a, b);
      a = uf_find(U, a);
      U[a].thr = e[i].w + (c / (float)U[a].size);
    }
  }
}
// full path compression variant (roots identical; only parent pointers differ)
template <int PF, int PF2>
void kruskal_full(const Edge12* e, size_t nE, size_t nV, float c, std::vector<UfElt>& u) {
  u.resize(nV);
  for (size_t i = 0; i < nV; ++i) { u[i].rank = 0; u[i].size = 1; u[i].p = (int)i; u[i].thr = c; }
  UfElt* U = u.data();
  auto find = [&](int x) { int y = x; while (y != U[y].p) y = U[y].p; while (U[x].p != y) { int n = U[x].p; U[x].p = y; x = n; } return y; };
  for (size_t i = 0; i < nE; ++i) {
    if (i + PF < nE) { __builtin_prefetch(U + e[i + PF].a); __builtin_prefetch(U + e[i + PF].b); }
    if (PF2 && i + PF2 < nE) { __builtin_prefetch(U + U[e[i + PF2].a].p); __builtin_prefetch(U + U[e[i + PF2].b].p); }
    int a = find(e[i].a), b = find(e[i].b);
    if (a != b && e[i].w <= U[a].thr && e[i].w <= U[b].thr) {
      uf_join(U, a, b);
      a = find(a);
      U[a].thr = e[i].w + (c / (float)U[a].size);
    }
  }
}

struct RootElt { int rank, size; float thr; int pad; };
template <int PF, int PF2>
void kruskal_soa(const Edge12* e, size_t nE, size_t nV, float c, std::vector<UfElt>& u) {
  std::vector<int> Pv(nV); std::vector<RootElt> Rv(nV);
  int* P = Pv.data(); RootElt* R = Rv.data();
  for (size_t i = 0; i < nV; ++i) { P[i] = (int)i; R[i].rank = 0; R[i].size = 1; R[i].thr = c; }
  auto find = [&](int x) { int y = x; while (y != P[y]) y = P[y]; while (P[x] != y) { int n = P[x]; P[x] = y; x = n; } return y; };
  for (size_t i = 0; i < nE; ++i) {
    if (i + PF < nE) { __builtin_prefetch(P + e[i + PF].a); __builtin_prefetch(P + e[i + PF].b); }
    if (PF2 && i + PF2 < nE) { const int pa = P[e[i + PF2].a], pb = P[e[i + PF2].b]; __builtin_prefetch(P + pa); __builtin_prefetch(P + pb); __builtin_prefetch(R + pa); __builtin_prefetch(R + pb); }
    int a = find(e[i].a), b = find(e[i].b);
    if (a != b && e[i].w <= R[a].thr && e[i].w <= R[b].thr) {
      int r;
      if (R[a].rank > R[b].rank) { P[b] = a; R[a].size += R[b].size; r = a; }
      else { P[a] = b; R[b].size += R[a].size; if (R[a].rank == R[b].rank) R[b].rank++; r = b; }
      R[r].thr = e[i].w + (c / (float)R[r].size);
    }
  }
  u.resize(nV);
  for (size_t i = 0; i < nV; ++i) { u[i].p = P[i]; u[i].rank = R[i].rank; u[i].size = R[i].size; u[i].thr = R[i].thr; }
}

template <int PF, int PF2>
void kruskal_walk(const Edge12* e, size_t nE, size_t nV, float c, std::vector<UfElt>& u) {
  u.resize(nV);
  for (size_t i = 0; i < nV; ++i) { u[i].rank = 0; u[i].size = 1; u[i].p = (int)i; u[i].thr = c; }
  UfElt* U = u.data();
  auto find = [&](int x) { int y = x; while (y != U[y].p) y = U[y].p; while (U[x].p != y) { int n = U[x].p; U[x].p = y; x = n; } return y; };
  for (size_t i = 0; i < nE; ++i) {
    if (i + PF < nE) { __builtin_prefetch(U + e[i + PF].a); __builtin_prefetch(U + e[i + PF].b); }
    if (i + PF2 < nE) { int y = e[i + PF2].a; y = U[y].p; y = U[y].p; __builtin_prefetch(U + y); y = e[i + PF2].b; y = U[y].p; y = U[y].p; __builtin_prefetch(U + y); }
    int a = find(e[i].a), b = find(e[i].b);
    if (a != b && e[i].w <= U[a].thr && e[i].w <= U[b].thr) {
      uf_join(U, a, b);
      a = find(a);
      U[a].thr = e[i].w + (c / (float)U[a].size);
    }
  }
}
int main() {
  FILE* f = fopen("edges.bin", "rb"); fseek(f, 0, SEEK_END); size_t n = ftell(f) / 12; fseek(f, 0, SEEK_SET);
  std::vector<Edge12> e(n); fread(e.data(), 12, n, f); fclose(f);
  size_t nV = 2000000;
  auto run = [&](const char* name, auto fn) {
    std::vector<UfElt> u; double best = 1e9; unsigned long h = 0;
    for (int r = 0; r < 3; ++r) { auto t0 = std::chrono::steady_clock::now(); fn(e.data(), n, nV, 0.01f, u); double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); if (ms < best) best = ms; }
    for (size_t q = 0; q < nV; ++q) { int y = q; while (y != u[y].p) y = u[y].p; h = h * 1000003u + y; }
    printf("%-24s %.1f ms  hash %lx\n", name, best, h);
  };
  run("pf24", kruskal<24, 0>);
  run("pf48", kruskal<48, 0>);
  run("pf96", kruskal<96, 0>);
  run("pf48+24", kruskal<48, 24>);
  run("pf96+48", kruskal<96, 48>);
  run("pf64+16", kruskal<64, 16>);
  run("full pf24", kruskal_full<24, 0>);
  run("full pf64+16", kruskal_full<64, 16>);
  run("walk 64+16", kruskal_walk<64, 16>);
  run("walk 96+32", kruskal_walk<96, 32>);
  run("walk 48+8", kruskal_walk<48, 8>);
  run("soa pf24", kruskal_soa<24, 0>);
  run("soa pf64+16", kruskal_soa<64, 16>);
  run("soa pf96+32", kruskal_soa<96, 32>);
  return 0;
}
