// mpcqp_band_host.hpp — host-side analysis of a CSC pattern for the sparse generic solve path (mpcqp_band.cuh): reverse
// Cuthill-McKee ordering of the KKT graph, bandwidth, band slots of every entry, row lists.  Plain C++ (also used by the
// host emulation of the kernel in tests/emul).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <queue>
#include <vector>

#include "mpcqp_band.cuh"

namespace mpcqp_band {

// ---- host-side analysis of the CSC pattern ------------------------------------------------------------------------------
struct HostPattern {
  int n = 0, m = 0, N = 0, w = 0, nnzP = 0, nnzA = 0;
  std::vector<int> Pc, Pi, Ac, Ai, Pr_ptr, Pr_pos, Pr_col, Ar_ptr, Ar_pos, Ar_col, slotP, slotA, perm, iperm;
};

// Reverse Cuthill-McKee on the graph of the KKT matrix (nodes 0..n-1 variables, n..N-1 constraints; edges: off-diagonal
// entries of P, entries of A).  Every connected component starts from a node of minimal degree found by repeated BFS
// (pseudo-peripheral node); neighbours are visited in order of increasing degree.
inline void rcm_order(int N, const std::vector<std::vector<int>>& adj, std::vector<int>* perm) {
  std::vector<int> deg(N), order; order.reserve(N);
  for (int i = 0; i < N; ++i) deg[i] = (int)adj[i].size();
  std::vector<char> seen(N, 0);
  std::vector<int> level(N);
  auto bfs_levels = [&](int root, std::vector<int>* nodes) {   // BFS of root's component restricted to unseen nodes; returns eccentricity
    nodes->clear();
    std::vector<int> mark;
    std::queue<int> q; q.push(root); level[root] = 0; mark.push_back(root);
    std::vector<char> local(N, 0); local[root] = 1;
    int ecc = 0;
    while (!q.empty()) {
      const int u = q.front(); q.pop(); nodes->push_back(u); ecc = level[u];
      for (int v : adj[u]) if (!seen[v] && !local[v]) { local[v] = 1; level[v] = level[u] + 1; q.push(v); }
    }
    return ecc;
  };
  std::vector<int> nodes;
  for (int s0 = 0; s0 < N; ++s0) {
    if (seen[s0]) continue;
    // pseudo-peripheral root of this component
    int root = s0, ecc = bfs_levels(root, &nodes);
    for (int it = 0; it < 8; ++it) {
      int best = -1;
      for (int u : nodes) if (level[u] == ecc && (best < 0 || deg[u] < deg[best])) best = u;
      if (best < 0 || best == root) break;
      std::vector<int> nodes2;
      const int e2 = bfs_levels(best, &nodes2);
      if (e2 <= ecc) { if (e2 == ecc && deg[best] < deg[root]) { root = best; nodes.swap(nodes2); } break; }
      root = best; ecc = e2; nodes.swap(nodes2);
    }
    // Cuthill-McKee from root
    std::queue<int> q; q.push(root); seen[root] = 1;
    while (!q.empty()) {
      const int u = q.front(); q.pop(); order.push_back(u);
      std::vector<int> nb;
      for (int v : adj[u]) if (!seen[v]) { seen[v] = 1; nb.push_back(v); }
      std::sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
      for (int v : nb) q.push(v);
    }
  }
  std::reverse(order.begin(), order.end());
  *perm = order;
}

// Returns false when the pattern is not eligible (bandwidth > kMaxBand): the caller then uses the dense kernel.
inline bool analyse(long long n, long long m, const int64_t* Pc, const int64_t* Pi, const int64_t* Ac, const int64_t* Ai, HostPattern* hp) {
  const int N = (int)(n + m);
  hp->n = (int)n; hp->m = (int)m; hp->N = N; hp->nnzP = (int)Pc[n]; hp->nnzA = (int)Ac[n];
  hp->Pc.assign(Pc, Pc + n + 1); hp->Pi.assign(Pi, Pi + Pc[n]); hp->Ac.assign(Ac, Ac + n + 1); hp->Ai.assign(Ai, Ai + Ac[n]);
  std::vector<std::vector<int>> adj((size_t)N);
  for (int j = 0; j < n; ++j) {
    for (long long t = Pc[j]; t < Pc[j + 1]; ++t) { const int i = (int)Pi[t]; if (i != j) { adj[(size_t)i].push_back(j); adj[(size_t)j].push_back(i); } }
    for (long long t = Ac[j]; t < Ac[j + 1]; ++t) { const int i = (int)(n + Ai[t]); adj[(size_t)i].push_back(j); adj[(size_t)j].push_back(i); }
  }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  rcm_order(N, adj, &hp->perm);
  hp->iperm.assign((size_t)N, 0);
  for (int r = 0; r < N; ++r) hp->iperm[(size_t)hp->perm[(size_t)r]] = r;
  int w = 0;
  for (int i = 0; i < N; ++i) for (int j : adj[(size_t)i]) w = std::max(w, std::abs(hp->iperm[(size_t)i] - hp->iperm[(size_t)j]));
  hp->w = w;
  if (w > kMaxBand) return false;
  const int W1 = w + 1;
  auto slot = [&](int oi, int oj) { int r = hp->iperm[(size_t)oi], c = hp->iperm[(size_t)oj]; if (r < c) std::swap(r, c); return r * W1 + (w - (r - c)); };
  hp->slotP.resize((size_t)hp->nnzP); hp->slotA.resize((size_t)hp->nnzA);
  // row lists: strictly upper part of P by row, A by row (ascending column: CSC columns are walked in order)
  std::vector<int> pr_cnt((size_t)n + 1, 0), ar_cnt((size_t)m + 1, 0);
  for (int j = 0; j < n; ++j) {
    for (long long t = Pc[j]; t < Pc[j + 1]; ++t) { hp->slotP[(size_t)t] = slot((int)Pi[t], j); if (Pi[t] != j) ++pr_cnt[(size_t)Pi[t] + 1]; }
    for (long long t = Ac[j]; t < Ac[j + 1]; ++t) { hp->slotA[(size_t)t] = slot((int)(n + Ai[t]), j); ++ar_cnt[(size_t)Ai[t] + 1]; }
  }
  for (int j = 0; j < n; ++j) pr_cnt[(size_t)j + 1] += pr_cnt[(size_t)j];
  for (int i = 0; i < m; ++i) ar_cnt[(size_t)i + 1] += ar_cnt[(size_t)i];
  hp->Pr_ptr = pr_cnt; hp->Ar_ptr = ar_cnt;
  hp->Pr_pos.resize((size_t)pr_cnt[(size_t)n]); hp->Pr_col.resize((size_t)pr_cnt[(size_t)n]);
  hp->Ar_pos.resize((size_t)ar_cnt[(size_t)m]); hp->Ar_col.resize((size_t)ar_cnt[(size_t)m]);
  std::vector<int> pf(pr_cnt.begin(), pr_cnt.end() - 1), af(ar_cnt.begin(), ar_cnt.end() - 1);
  for (int j = 0; j < n; ++j) {
    for (long long t = Pc[j]; t < Pc[j + 1]; ++t) if (Pi[t] != j) { const int r = (int)Pi[t]; hp->Pr_pos[(size_t)pf[(size_t)r]] = (int)t; hp->Pr_col[(size_t)pf[(size_t)r]++] = j; }
    for (long long t = Ac[j]; t < Ac[j + 1]; ++t) { const int r = (int)Ai[t]; hp->Ar_pos[(size_t)af[(size_t)r]] = (int)t; hp->Ar_col[(size_t)af[(size_t)r]++] = j; }
  }
  return true;
}

// Flattened int32 image of the pattern for one upload; offsets in `off` (same order as the Pattern pointers).
inline int pattern_build(long long n, long long m, const int64_t* Pc, const int64_t* Pi, const int64_t* Ac, const int64_t* Ai, std::vector<int>* flat, int* off, int* w) {
  HostPattern hp;
  if (n + m > 8192 || !analyse(n, m, Pc, Pi, Ac, Ai, &hp)) { *w = hp.w; return 0; }
  const std::vector<int>* parts[] = { &hp.Pc, &hp.Pi, &hp.Ac, &hp.Ai, &hp.Pr_ptr, &hp.Pr_pos, &hp.Pr_col, &hp.Ar_ptr, &hp.Ar_pos, &hp.Ar_col, &hp.slotP, &hp.slotA, &hp.perm, &hp.iperm };
  flat->clear();
  for (int k = 0; k < 14; ++k) { off[k] = (int)flat->size(); flat->insert(flat->end(), parts[k]->begin(), parts[k]->end()); flat->push_back(0); }
  *w = hp.w;
  return 1;
}

}  // namespace mpcqp_band
