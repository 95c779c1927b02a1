// symbolic.cpp -- nested-dissection ordering + assembly tree + front structures (host).
// See symbolic.h.  Pure index work; runs once per graph structure.
#include "symbolic.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <condition_variable>
#include <mutex>
#include <set>
#include <thread>

namespace {

struct NdNode {
  std::vector<int> verts;  // pivots of this front (block indices)
  std::vector<int> kids;   // nested-dissection children (node ids)
};

struct Nd {
  int nb;
  const int* dim;
  std::vector<int> xadj, adj;
  std::vector<int> region_store;  // current region label of every vertex
  std::vector<int> lvl_store;     // BFS level scratch
  // views used by the algorithm: a sub-dissection running on a pool thread shares the arrays of its
  // parent (its vertices, and the labels it hands out, are disjoint from everybody else's)
  const int* xa = nullptr;
  const int* ad = nullptr;
  int* region = nullptr;
  int* lvl = nullptr;
  std::vector<NdNode> nodes;
  int leaf_size;
  int next_label = 1;
  // recursion below `defer_depth` is not run but recorded (vertex set, label, parent node): the
  // caller runs those sub-dissections in parallel and splices their trees in, in this order
  struct Deferred { std::vector<int> verts; int lab; int parent; int hint; };
  int defer_depth = -1;
  std::vector<Deferred> deferred;

  // BFS inside the current region from r; fills order (visit order) and lvl[]; returns the number of levels.
  // Invariant (kept by order_region): lvl[v] == -1 exactly for the not yet visited vertices of the region being
  // searched; every other vertex a search can reach -- the separators around the region -- carries lvl >= 0, so the
  // neighbour test is ONE random read instead of a region label and a level.  With `weights` the search also
  // accumulates, per level, the weight of the level (w), of its vertices with a neighbour in the next level (wnext)
  // and in the previous one (wprev): when v is scanned all of level(v) - 1 and the part of level(v) + 1 reached so
  // far are labelled and the rest of level(v) + 1 is labelled by this very scan, so nothing is missed -- the separate
  // sweep over the adjacency that used to compute them is gone.
  template <bool WEIGHTS>
  int bfs(int r, std::vector<int>& order, std::vector<int>& level_start, std::vector<double>* w = nullptr,
          std::vector<double>* wnext = nullptr, std::vector<double>* wprev = nullptr) {
    order.clear();
    level_start.clear();
    order.push_back(r);
    lvl[r] = 0;
    level_start.push_back(0);
    if (WEIGHTS) { w->assign(1, 0.0); wnext->assign(1, 0.0); wprev->assign(1, 0.0); }
    size_t head = 0;
    int cur = 0;
    while (head < order.size()) {
      const int v = order[head];
      const int lv = lvl[v];
      if (lv != cur) {
        cur = lv;
        level_start.push_back((int)head);
        if (WEIGHTS) { w->push_back(0.0); wnext->push_back(0.0); wprev->push_back(0.0); }
      }
      head++;
      bool hn = false, hp = false;
      for (int p = xa[v]; p < xa[v + 1]; p++) {
        const int u = ad[p];
        const int lu = lvl[u];
        if (lu < 0) {
          lvl[u] = lv + 1;
          order.push_back(u);
          hn = true;
        } else if (WEIGHTS) {
          if (lu == lv + 1) hn = true;
          else if (lu == lv - 1) hp = true;
        }
      }
      if (WEIGHTS) {
        (*w)[cur] += dim[v];
        if (hn) (*wnext)[cur] += dim[v];
        if (hp) (*wprev)[cur] += dim[v];
      }
    }
    level_start.push_back((int)order.size());
    return (int)level_start.size() - 1;
  }

  // orders the vertex set `verts` (all carrying region label `lab`); appends the resulting
  // subtree roots to `roots`.  `hint`: a vertex of `verts` known to lie at an end of its component (the parent
  // bisection's own root for the near half, the vertex its search reached last for the far half): the search from
  // it is taken as the level structure at once -- one sweep over the adjacency per bisection instead of two.
  void order_region(std::vector<int>& verts, int lab, std::vector<int>& roots, int depth = 0, int hint = -1) {
    for (int v : verts) lvl[v] = -1;
    std::vector<int> order, level_start;
    std::vector<double> w, wnext, wprev;
    auto leaf = [&](const std::vector<int>& comp) {
      NdNode nd;
      nd.verts = comp;
      nodes.push_back(nd);
      roots.push_back((int)nodes.size() - 1);
    };
    // MEASURED AND OFF BY DEFAULT (SLAM_B200_ND_HINT=1): the hinted search saves a sweep per bisection (nested
    // dissection of the 10-lap graph 4.4 -> 2.9 ms on one host thread) but its level structures are a little worse --
    // assembly trees of 13 instead of 10 levels on the 1- and 10-lap graphs, largest corridor front 95 instead of 72
    // rows -- and every level is a dependent launch in each of the 10 iterations.
    static const bool use_hint = getenv("SLAM_B200_ND_HINT") != nullptr;
    if (hint >= 0 && use_hint) {
      const int nl = bfs<true>(hint, order, level_start, &w, &wnext, &wprev);
      if ((int)order.size() <= leaf_size) leaf(order);
      else dissect(order, level_start, w, wnext, wprev, nl, lab, roots, depth);
    }
    // the (other) connected components
    std::vector<std::vector<int>> comps;
    for (int v : verts) {
      if (lvl[v] >= 0) continue;
      bfs<false>(v, order, level_start);
      comps.push_back(order);
    }
    for (auto& comp : comps) {
      if ((int)comp.size() <= leaf_size) {
        leaf(comp);
        continue;
      }
      // pseudo-peripheral start: the breadth-first search that found the component (comp is its visit order from
      // comp[0]) is the first sweep; the level structure is the second sweep's, rooted at the vertex the first one
      // reached last.  (A third sweep from the second one's last vertex was measured: same fill on the trackdrive
      // and corridor graphs to within 1 %, one more pass over the adjacency per bisection.)
      const int r = comp.back();
      for (int v : comp) lvl[v] = -1;
      const int nl = bfs<true>(r, order, level_start, &w, &wnext, &wprev);
      dissect(order, level_start, w, wnext, wprev, nl, lab, roots, depth);
    }
  }

  // bisects one connected component given its level structure (order = visit order, lvl[] = levels, per-level weights)
  void dissect(const std::vector<int>& order, const std::vector<int>& level_start, const std::vector<double>& w,
               const std::vector<double>& wnext, const std::vector<double>& wprev, int nl, int lab, std::vector<int>& roots,
               int depth) {
    (void)level_start;
    {
      const std::vector<int>& comp = order;
      double W = 0;
      for (double x : w) W += x;
      // candidates: cut between level l and l+1; option 1 separator = boundary of level l
      // (towards l+1), option 2 = boundary of level l+1 (towards l)
      int bestL = -1, bestOpt = 0;
      double bestCost = 1e300;
      double below = 0;
      for (int l = 0; l + 1 < nl; l++) {
        below += w[l];
        for (int opt = 1; opt <= 2; opt++) {
          double sep = opt == 1 ? wnext[l] : wprev[l + 1];
          double a = opt == 1 ? below - sep : below;
          double b = opt == 1 ? W - below : W - below - sep;
          if (a <= 0 || b <= 0) continue;
          double bal = std::min(a, b) / (a + b);
          if (bal < 0.2) continue;
          double cost = sep * (1.0 + 0.5 * (0.5 - bal));
          if (cost < bestCost) {
            bestCost = cost;
            bestL = l;
            bestOpt = opt;
          }
        }
      }
      if (bestL < 0) {
        // no balanced level cut (e.g. a near-clique): relax balance, else dense leaf
        below = 0;
        for (int l = 0; l + 1 < nl; l++) {
          below += w[l];
          for (int opt = 1; opt <= 2; opt++) {
            double sep = opt == 1 ? wnext[l] : wprev[l + 1];
            double a = opt == 1 ? below - sep : below;
            double b = opt == 1 ? W - below : W - below - sep;
            if (a <= 0 || b <= 0) continue;
            double cost = sep / std::min(a, b);
            if (sep < 0.5 * W && cost < bestCost) {
              bestCost = cost;
              bestL = l;
              bestOpt = opt;
            }
          }
        }
      }
      if (bestL < 0) {
        NdNode nd;
        nd.verts = comp;
        nodes.push_back(nd);
        roots.push_back((int)nodes.size() - 1);
        return;
      }
      int sepLevel = bestOpt == 1 ? bestL : bestL + 1;
      int other = bestOpt == 1 ? bestL + 1 : bestL;
      std::vector<int> S, A, B;
      for (int q = 0; q < (int)order.size(); q++) {
        int v = order[q];
        int l = lvl[v];
        bool inSep = false;
        if (l == sepLevel) {
          for (int p = xa[v]; p < xa[v + 1]; p++) {
            int u = ad[p];
            if (region[u] == lab && lvl[u] == other) { inSep = true; break; }
          }
        }
        if (inSep) S.push_back(v);
        else if (l <= bestL) A.push_back(v);
        else B.push_back(v);
      }
      int la = next_label++, lb = next_label++;
      for (int v : A) region[v] = la;
      for (int v : B) region[v] = lb;
      for (int v : S) { region[v] = 0; lvl[v] = 0x3fffffff; }  // separators leave every region: visited, on no level
      NdNode nd;
      nd.verts = S;
      nodes.push_back(nd);
      int me = (int)nodes.size() - 1;
      std::vector<int> kids;
      const int hintA = A.empty() ? -1 : A.front(), hintB = B.empty() ? -1 : B.back();
      if (depth == defer_depth) {
        deferred.push_back({std::move(A), la, me, hintA});
        deferred.push_back({std::move(B), lb, me, hintB});
      } else {
        order_region(A, la, kids, depth + 1, hintA);
        order_region(B, lb, kids, depth + 1, hintB);
      }
      nodes[me].kids = kids;
      roots.push_back(me);
    }
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// Constrained minimum degree on the block graph.  rank[v] orders groups (all vertices of a lower
// rank are eliminated before any vertex of a higher rank: nested-dissection leaves first, then the
// separators bottom-up); inside a group the vertex of minimum weighted external degree goes next.
// Hub vertices (a landmark seen from hundreds of poses) therefore stay until their spokes are gone,
// which is what keeps fill low on pose-landmark graphs.  Exact elimination graph, sorted adjacency.
// ------------------------------------------------------------------------------------------------

// A small persistent pool of host threads (created once per process: thread creation costs
// milliseconds inside container sandboxes, which would eat the whole gain of a per-call pool).
class HostPool {
 public:
  static HostPool& get() {
    static HostPool p;
    return p;
  }
  int size() const { return (int)workers_.size() + 1; }
  // runs fn(0..njobs-1), the caller participates; returns when every job is done
  void run(int njobs, const std::function<void(int)>& fn) {
    if (njobs <= 0) return;
    if (workers_.empty() || njobs == 1) {
      for (int j = 0; j < njobs; j++) fn(j);
      return;
    }
    // one parallel section at a time: contexts used from different host threads share this pool
    std::lock_guard<std::mutex> one_run(run_m_);
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      njobs_ = njobs;
      next_ = 0;
      pending_ = njobs;
      generation_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
    njobs_ = 0;
  }

 private:
  HostPool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)std::min<unsigned>(std::max(1u, hw), 16u);
    if (const char* e = getenv("SLAM_B200_SYM_THREADS")) n = std::max(1, atoi(e));
    for (int t = 1; t < n; t++) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      generation_++;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  // Jobs are claimed under the mutex, together with the function they belong to: a worker that is
  // still on its way out of the previous run when the next one is being set up either sees the old
  // run exhausted or a complete new one -- never the new job count against the old counter (which
  // would run a job twice and let run() return early).  Jobs are coarse (a region, a slice of the
  // fronts), the lock is not on any hot path.
  void work() {
    std::unique_lock<std::mutex> lk(m_);
    while (next_ < njobs_) {
      const int j = next_++;
      const std::function<void(int)>* fn = fn_;
      lk.unlock();
      (*fn)(j);
      lk.lock();
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, run_m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int njobs_ = 0, pending_ = 0, next_ = 0;  // all guarded by m_
  unsigned long generation_ = 0;
  bool stop_ = false;
};

// Generic kernel: eliminates every vertex with rank <= max_rank in (rank, weighted degree) order on
// the elimination graph `adj` (sorted adjacency lists, updated in place: what is left afterwards is
// the elimination graph of the remaining vertices).  Appends to `order`.
static void min_degree_eliminate(std::vector<std::vector<int>>& adj, const int* dim, const std::vector<int>& rank,
                                 int max_rank, std::vector<int>& order) {
  const int nb = (int)adj.size();
  int nrank = 1;
  for (int v = 0; v < nb; v++)
    if (rank[v] <= max_rank) nrank = std::max(nrank, rank[v] + 1);
  std::vector<long> wdeg(nb, 0);
  std::vector<std::vector<int>> head(nrank);
  std::vector<int> nxt(nb, -1), prv(nb, -1), remaining(nrank, 0), mindeg(nrank, 0);
  auto in_play = [&](int v) { return rank[v] <= max_rank; };
  auto bucket_insert = [&](int v) {
    std::vector<int>& h = head[rank[v]];
    long d = wdeg[v];
    if ((long)h.size() <= d) h.resize(d + 1 + (d >> 2), -1);
    nxt[v] = h[d];
    prv[v] = -1;
    if (h[d] >= 0) prv[h[d]] = v;
    h[d] = v;
    if (d < mindeg[rank[v]]) mindeg[rank[v]] = (int)d;
  };
  auto bucket_remove = [&](int v) {
    std::vector<int>& h = head[rank[v]];
    if (prv[v] >= 0) nxt[prv[v]] = nxt[v];
    else h[wdeg[v]] = nxt[v];
    if (nxt[v] >= 0) prv[nxt[v]] = prv[v];
  };
  for (int v = 0; v < nb; v++) {
    for (int u : adj[v]) wdeg[v] += dim[u];
    if (in_play(v)) remaining[rank[v]]++;
  }
  for (int v = nb - 1; v >= 0; v--)
    if (in_play(v)) bucket_insert(v);  // ties resolved towards the lower index
  std::vector<int> N;
  for (int rk = 0; rk < nrank; rk++) {
    while (remaining[rk] > 0) {
      std::vector<int>& h = head[rk];
      int d = mindeg[rk];
      while (d < (int)h.size() && h[d] < 0) d++;
      mindeg[rk] = d;
      const int v = h[d];
      bucket_remove(v);
      remaining[rk]--;
      order.push_back(v);
      N.swap(adj[v]);
      adj[v].clear();
      adj[v].shrink_to_fit();
      for (int u : N) {
        const bool play = in_play(u);
        if (play) bucket_remove(u);
        std::vector<int>& A = adj[u];
        long w = wdeg[u];
        auto it = std::lower_bound(A.begin(), A.end(), v);
        if (it != A.end() && *it == v) { A.erase(it); w -= dim[v]; }
        // insert the members of N that are missing (usually none: the neighbours of a low-degree
        // vertex mostly know each other already)
        for (int x : N) {
          if (x == u) continue;
          auto jt = std::lower_bound(A.begin(), A.end(), x);
          if (jt == A.end() || *jt != x) { A.insert(jt, x); w += dim[x]; }
        }
        wdeg[u] = w;
        if (play) bucket_insert(u);
      }
    }
  }
}

// The same elimination on a dense bit-matrix adjacency (one row of ceil(nb/64) words per vertex),
// for graphs of at most a few thousand vertices -- every nested-dissection region with its halo and
// the separator graph are that small.  Merging the pivot's neighbourhood into a neighbour's row is a
// word-wise OR, and only the bits that are new cost anything beyond that, so an elimination is
// O(|N| * nb/64) instead of O(|N|^2 log) sorted-vector searches.  Same buckets, same visiting order
// (ascending neighbours) => the very same elimination order as min_degree_eliminate.
static void min_degree_eliminate_dense(std::vector<std::vector<int>>& adj, const int* dim, const std::vector<int>& rank,
                                       int max_rank, std::vector<int>& order) {
  const int nb = (int)adj.size();
  const int W = (nb + 63) >> 6;
  std::vector<uint64_t> bits((size_t)nb * W, 0);
  int nrank = 1;
  for (int v = 0; v < nb; v++)
    if (rank[v] <= max_rank) nrank = std::max(nrank, rank[v] + 1);
  std::vector<long> wdeg(nb, 0);
  std::vector<std::vector<int>> head(nrank);
  std::vector<int> nxt(nb, -1), prv(nb, -1), remaining(nrank, 0), mindeg(nrank, 0);
  std::vector<char> gone(nb, 0);
  auto in_play = [&](int v) { return rank[v] <= max_rank; };
  auto bucket_insert = [&](int v) {
    std::vector<int>& h = head[rank[v]];
    long d = wdeg[v];
    if ((long)h.size() <= d) h.resize(d + 1 + (d >> 2), -1);
    nxt[v] = h[d];
    prv[v] = -1;
    if (h[d] >= 0) prv[h[d]] = v;
    h[d] = v;
    if (d < mindeg[rank[v]]) mindeg[rank[v]] = (int)d;
  };
  auto bucket_remove = [&](int v) {
    std::vector<int>& h = head[rank[v]];
    if (prv[v] >= 0) nxt[prv[v]] = nxt[v];
    else h[wdeg[v]] = nxt[v];
    if (nxt[v] >= 0) prv[nxt[v]] = prv[v];
  };
  for (int v = 0; v < nb; v++) {
    uint64_t* row = bits.data() + (size_t)v * W;
    for (int u : adj[v]) {
      wdeg[v] += dim[u];
      row[u >> 6] |= 1ull << (u & 63);
    }
    if (in_play(v)) remaining[rank[v]]++;
  }
  for (int v = nb - 1; v >= 0; v--)
    if (in_play(v)) bucket_insert(v);  // ties resolved towards the lower index
  std::vector<int> N;
  std::vector<uint64_t> Nb(W);
  for (int rk = 0; rk < nrank; rk++) {
    while (remaining[rk] > 0) {
      std::vector<int>& h = head[rk];
      int d = mindeg[rk];
      while (d < (int)h.size() && h[d] < 0) d++;
      mindeg[rk] = d;
      const int v = h[d];
      bucket_remove(v);
      remaining[rk]--;
      order.push_back(v);
      gone[v] = 1;
      uint64_t* rv = bits.data() + (size_t)v * W;
      N.clear();
      for (int k = 0; k < W; k++) {
        uint64_t w = Nb[k] = rv[k];
        rv[k] = 0;
        while (w) {
          N.push_back((k << 6) + __builtin_ctzll(w));
          w &= w - 1;
        }
      }
      const int vk = v >> 6;
      const uint64_t vbit = 1ull << (v & 63);
      for (int u : N) {
        const bool play = in_play(u);
        if (play) bucket_remove(u);
        uint64_t* ru = bits.data() + (size_t)u * W;
        long w = wdeg[u];
        if (ru[vk] & vbit) { ru[vk] &= ~vbit; w -= dim[v]; }
        const int uk = u >> 6;
        const uint64_t ubit = 1ull << (u & 63);
        for (int k = 0; k < W; k++) {
          uint64_t nw = Nb[k] & ~ru[k];
          if (k == uk) nw &= ~ubit;
          if (!nw) continue;
          ru[k] |= nw;
          while (nw) {
            w += dim[(k << 6) + __builtin_ctzll(nw)];
            nw &= nw - 1;
          }
        }
        wdeg[u] = w;
        if (play) bucket_insert(u);
      }
    }
  }
  // hand the elimination graph of the remaining vertices back as sorted lists
  for (int v = 0; v < nb; v++) {
    adj[v].clear();
    if (gone[v]) { adj[v].shrink_to_fit(); continue; }
    const uint64_t* row = bits.data() + (size_t)v * W;
    for (int k = 0; k < W; k++) {
      uint64_t w = row[k];
      while (w) {
        adj[v].push_back((k << 6) + __builtin_ctzll(w));
        w &= w - 1;
      }
    }
  }
}

constexpr int MD_DENSE_LIMIT = 4096;  // vertices; a 4096 x 4096 bit matrix is 2 MB

static void min_degree_dispatch(std::vector<std::vector<int>>& adj, const int* dim, const std::vector<int>& rank,
                                int max_rank, std::vector<int>& order) {
  if ((int)adj.size() <= MD_DENSE_LIMIT && !getenv("SLAM_B200_MD_SPARSE")) min_degree_eliminate_dense(adj, dim, rank, max_rank, order);
  else min_degree_eliminate(adj, dim, rank, max_rank, order);
}

// Constrained minimum degree.  The interiors of the nested-dissection regions (rank 0) do not touch
// each other -- only the separators around them -- so every region is ordered independently on its
// own copy of (interior + halo) by a pool of host threads; the fill each region leaves among its
// halo vertices is merged into the separator graph, which is then ordered rank by rank.
static std::vector<int> constrained_min_degree(int nb, const int* dim, const std::vector<int>& xadj,
                                               const std::vector<int>& adjv, const std::vector<int>& rank,
                                               const std::vector<std::vector<int>>& regions) {
  auto tdbg0 = std::chrono::steady_clock::now();
  std::vector<int> order;
  order.reserve(nb);
  const int nreg = (int)regions.size();
  std::vector<std::vector<int>> reg_order(nreg);
  std::vector<std::vector<std::pair<int, int>>> reg_fill(nreg);  // halo-halo fill edges (global ids)
  auto do_region = [&](int ri) {
    auto tr0 = std::chrono::steady_clock::now();
    const std::vector<int>& R = regions[ri];
    // local numbering: interior first (region order), then halo (ascending global id); the
    // global -> local map is a per-thread scratch array that is reset entry by entry afterwards
    static thread_local std::vector<int> g2l;
    if ((int)g2l.size() < nb) g2l.assign(nb, -1);
    std::vector<int> loc2g(R);
    const int nint = (int)R.size();
    for (int k = 0; k < nint; k++) g2l[R[k]] = k;
    auto find_local = [&](int gv) -> int { return g2l[gv]; };
    std::vector<int> halo;
    for (int v : R)
      for (int p = xadj[v]; p < xadj[v + 1]; p++) {
        int u = adjv[p];
        if (g2l[u] == -1) { g2l[u] = -2; halo.push_back(u); }
      }
    std::sort(halo.begin(), halo.end());
    for (int hv : halo) { g2l[hv] = (int)loc2g.size(); loc2g.push_back(hv); }
    const int nloc = (int)loc2g.size();
    std::vector<std::vector<int>> ladj(nloc);
    std::vector<int> ldim(nloc), lrank(nloc);
    for (int k = 0; k < nloc; k++) { ldim[k] = dim[loc2g[k]]; lrank[k] = k < nint ? 0 : 1; }
    for (int k = 0; k < nint; k++) {
      int v = loc2g[k];
      for (int p = xadj[v]; p < xadj[v + 1]; p++) {
        int lu = find_local(adjv[p]);
        ladj[k].push_back(lu);
        if (lu >= nint) ladj[lu].push_back(k);  // halo side of an interior-halo edge
      }
    }
    for (auto& a : ladj) {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
    std::vector<int> lord;
    lord.reserve(nint);
    min_degree_dispatch(ladj, ldim.data(), lrank, 0, lord);
    reg_order[ri].reserve(nint);
    for (int k : lord) reg_order[ri].push_back(loc2g[k]);
    for (int k = nint; k < nloc; k++)
      for (int u : ladj[k])
        if (u > k) reg_fill[ri].push_back({loc2g[k], loc2g[u]});
    for (int gv : loc2g) g2l[gv] = -1;
    if (getenv("SLAM_B200_SYM_DEBUG"))
      fprintf(stderr, "[symbolic] region %d: %d interior, %d halo, start %.4f dur %.4f s\n", ri, nint, nloc - nint, std::chrono::duration<double>(tr0 - tdbg0).count(), std::chrono::duration<double>(std::chrono::steady_clock::now() - tr0).count());
  };
  {
    // largest regions first: better balance when a few regions dominate
    std::vector<int> by_size(nreg);
    std::iota(by_size.begin(), by_size.end(), 0);
    std::sort(by_size.begin(), by_size.end(), [&](int a, int b) { return regions[a].size() > regions[b].size(); });
    HostPool::get().run(nreg, [&](int j) { do_region(by_size[j]); });
  }
  if (getenv("SLAM_B200_SYM_DEBUG"))
    fprintf(stderr, "[symbolic] regions joined %.4f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tdbg0).count());
  if (getenv("SLAM_B200_SYM_DEBUG")) {
    size_t mx = 0, tot = 0, fill = 0;
    for (int ri = 0; ri < nreg; ri++) { mx = std::max(mx, regions[ri].size()); tot += regions[ri].size(); fill += reg_fill[ri].size(); }
    fprintf(stderr, "[symbolic] %d regions, max %zu, total %zu of %d, halo fill edges %zu\n", nreg, mx, tot, nb, fill);
  }
  for (int ri = 0; ri < nreg; ri++) order.insert(order.end(), reg_order[ri].begin(), reg_order[ri].end());
  // separator graph: original edges among the remaining vertices + the regions' fill
  std::vector<char> gone(nb, 0);
  for (int v : order) gone[v] = 1;
  std::vector<std::vector<int>> adj(nb);
  for (int v = 0; v < nb; v++) {
    if (gone[v]) continue;
    for (int p = xadj[v]; p < xadj[v + 1]; p++)
      if (!gone[adjv[p]]) adj[v].push_back(adjv[p]);
  }
  for (int ri = 0; ri < nreg; ri++)
    for (auto& e : reg_fill[ri]) { adj[e.first].push_back(e.second); adj[e.second].push_back(e.first); }
  for (int v = 0; v < nb; v++) {
    std::sort(adj[v].begin(), adj[v].end());
    adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end());
  }
  // the separator graph on its own compact numbering (ascending global id, so ties break the same
  // way as on the full numbering)
  int maxr = 0;
  for (int v = 0; v < nb; v++) maxr = std::max(maxr, rank[v]);
  std::vector<int> loc(nb, -1), glob;
  for (int v = 0; v < nb; v++)
    if (!gone[v]) { loc[v] = (int)glob.size(); glob.push_back(v); }
  const int ns = (int)glob.size();
  std::vector<std::vector<int>> sadj(ns);
  std::vector<int> sdim(ns), srank(ns);
  for (int k = 0; k < ns; k++) {
    const int v = glob[k];
    sdim[k] = dim[v];
    srank[k] = rank[v];
    sadj[k].reserve(adj[v].size());
    for (int u : adj[v]) sadj[k].push_back(loc[u]);  // ascending: loc is monotone
  }
  if (getenv("SLAM_B200_SYM_DEBUG"))
    fprintf(stderr, "[symbolic] regions+merge %.4f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tdbg0).count());
  std::vector<int> rest;
  rest.reserve(ns);
  min_degree_dispatch(sadj, sdim.data(), srank, maxr, rest);
  for (int& k : rest) k = glob[k];
  if (getenv("SLAM_B200_SYM_DEBUG"))
    fprintf(stderr, "[symbolic] +separators %.4f s (%zu vertices)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tdbg0).count(), rest.size());
  order.insert(order.end(), rest.begin(), rest.end());
  return order;
}

void symbolic_analyze(int nb, const int* dim, int nnb, const int* off_a, const int* off_b,
                      const int* hoff_diag, const int* hoff_off, int leaf_size, Symbolic& S, int amalgamate,
                      const std::function<void(bool)>* structure_ready) {
  struct ReadyOnce {  // signals the caller exactly once, also on the way out of an exception
    const std::function<void(bool)>* fn;
    void fire(bool ok) { if (fn) { const std::function<void(bool)>* f = fn; fn = nullptr; (*f)(ok); } }
    ~ReadyOnce() { fire(false); }
  } ready{structure_ready};
  auto t0 = std::chrono::steady_clock::now();
  S = Symbolic();
  S.nb = nb;
  S.dim.assign(dim, dim + nb);
  // ---- adjacency ----
  Nd nd;
  nd.nb = nb;
  nd.dim = dim;
  nd.leaf_size = std::max(1, leaf_size);
  nd.xadj.assign(nb + 1, 0);
  for (int k = 0; k < nnb; k++) {
    nd.xadj[off_a[k] + 1]++;
    nd.xadj[off_b[k] + 1]++;
  }
  for (int v = 0; v < nb; v++) nd.xadj[v + 1] += nd.xadj[v];
  nd.adj.assign(nd.xadj[nb], 0);
  {
    std::vector<int> cur(nd.xadj.begin(), nd.xadj.end() - 1);
    for (int k = 0; k < nnb; k++) {
      nd.adj[cur[off_a[k]]++] = off_b[k];
      nd.adj[cur[off_b[k]]++] = off_a[k];
    }
  }
  // ---- stage 1: coarse nested dissection (regions of <= leaf_size vertices) gives the group
  // ranks: all region interiors first, then separators from the deepest level up to the root
  nd.region_store.assign(nb, 1);
  nd.lvl_store.assign(nb, -1);
  nd.xa = nd.xadj.data();
  nd.ad = nd.adj.data();
  nd.region = nd.region_store.data();
  nd.lvl = nd.lvl_store.data();
  nd.next_label = 2;
  std::vector<int> all(nb), roots;
  std::iota(all.begin(), all.end(), 0);
  // Only the top bisection runs here; its two halves are bisected side by side on the pool threads, and the (up
  // to four) quarters below them are dissected side by side after that.  The split depths are fixed and every part
  // gets its own label range from its position, so the result does not depend on the thread count.
  nd.defer_depth = nb >= 4096 ? 0 : -1;
  nd.order_region(all, 1, roots, 0);
  if (getenv("SLAM_B200_SYM_DEBUG"))
    fprintf(stderr, "[symbolic] nd top level done at %.4f s, %zu deferred parts\n",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), nd.deferred.size());
  if (!nd.deferred.empty()) {
    auto init_sub = [&](Nd& q, int label_block, int defer) {
      q.nb = nb; q.dim = dim; q.leaf_size = nd.leaf_size;
      q.xa = nd.xa; q.ad = nd.ad; q.region = nd.region; q.lvl = nd.lvl;
      q.next_label = 2 + 2 * nb * label_block;  // label ranges never overlap (a dissection of n vertices hands out < 2n labels)
      q.defer_depth = defer;
    };
    auto splice = [](Nd& into, std::vector<Nd>& subs, std::vector<std::vector<int>>& sub_roots) {
      for (size_t t = 0; t < subs.size(); t++) {
        const int off = (int)into.nodes.size();
        for (NdNode& node : subs[t].nodes) {
          for (int& k : node.kids) k += off;
          into.nodes.push_back(std::move(node));
        }
        for (int rt : sub_roots[t]) into.nodes[into.deferred[t].parent].kids.push_back(rt + off);
      }
      into.deferred.clear();
    };
    // halves: one more bisection each, their own parts deferred again
    const int n1 = (int)nd.deferred.size();
    std::vector<Nd> half(n1);
    std::vector<std::vector<int>> half_roots(n1);
    HostPool::get().run(n1, [&](int t) {
      init_sub(half[t], 1 + t, 0);
      half[t].order_region(nd.deferred[t].verts, nd.deferred[t].lab, half_roots[t], 0, nd.deferred[t].hint);
    });
    // quarters of all halves, flattened
    std::vector<std::pair<int, int>> parts;
    for (int h = 0; h < n1; h++)
      for (int k = 0; k < (int)half[h].deferred.size(); k++) parts.push_back({h, k});
    const int n2 = (int)parts.size();
    std::vector<Nd> quarter(n2);
    std::vector<std::vector<int>> quarter_roots(n2);
    HostPool::get().run(n2, [&](int t) {
      Nd::Deferred& d = half[parts[t].first].deferred[parts[t].second];
      init_sub(quarter[t], 1 + n1 + t, -1);
      quarter[t].order_region(d.verts, d.lab, quarter_roots[t], 0, d.hint);
    });
    for (int h = 0, t0q = 0; h < n1; h++) {
      const int cnt = (int)half[h].deferred.size();
      std::vector<Nd> subs(std::make_move_iterator(quarter.begin() + t0q), std::make_move_iterator(quarter.begin() + t0q + cnt));
      std::vector<std::vector<int>> sr(quarter_roots.begin() + t0q, quarter_roots.begin() + t0q + cnt);
      splice(half[h], subs, sr);
      t0q += cnt;
    }
    splice(nd, half, half_roots);
  }
  if (getenv("SLAM_B200_SYM_DEBUG"))
    fprintf(stderr, "[symbolic] nd parts done at %.4f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  std::vector<int> rank(nb, 0);
  {
    // depth of every nested-dissection node; leaves (no kids) get rank 0, a separator at depth d
    // gets rank maxdepth - d + 1
    std::vector<int> depth(nd.nodes.size(), 0);
    int maxdepth = 0;
    std::vector<int> stack(roots.begin(), roots.end());
    while (!stack.empty()) {
      int f = stack.back();
      stack.pop_back();
      maxdepth = std::max(maxdepth, depth[f]);
      for (int k : nd.nodes[f].kids) { depth[k] = depth[f] + 1; stack.push_back(k); }
    }
    for (size_t f = 0; f < nd.nodes.size(); f++) {
      int r = nd.nodes[f].kids.empty() ? 0 : (maxdepth - depth[f] + 1);
      for (int v : nd.nodes[f].verts) rank[v] = r;
    }
  }
  S.t_nd = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // ---- stage 2: constrained minimum degree -> elimination order ----
  std::vector<std::vector<int>> regions;  // interiors of the nested-dissection leaves (rank 0)
  for (auto& node : nd.nodes)
    if (node.kids.empty() && !node.verts.empty()) regions.push_back(node.verts);
  std::vector<int> order = constrained_min_degree(nb, dim, nd.xadj, nd.adj, rank, regions);
  S.t_md = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - S.t_nd;
  std::vector<int> epos(nb);
  for (int k = 0; k < nb; k++) epos[order[k]] = k;
  // ---- stage 3: block elimination tree + column structures, postorder, supernodes ----
  auto tdb = std::chrono::steady_clock::now();
  auto dbg = [&](const char* what) {
    if (!getenv("SLAM_B200_SYM_DEBUG")) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[symbolic] stage3 %-24s %.4f s\n", what, std::chrono::duration<double>(now - tdb).count());
    tdb = now;
  };
  // Elimination tree first (Liu's algorithm with path compression -- no column structures needed),
  // then its postorder, and only then the column structures, once, in the final numbering.
  std::vector<int> eparent(nb, -1);
  {
    std::vector<int> anc(nb, -1);
    for (int k = 0; k < nb; k++) {
      const int v = order[k];
      for (int p = nd.xadj[v]; p < nd.xadj[v + 1]; p++) {
        int j = epos[nd.adj[p]];
        if (j >= k) continue;
        while (anc[j] != -1 && anc[j] != k) {  // climb to the root of j's subtree, compressing
          const int t = anc[j];
          anc[j] = k;
          j = t;
        }
        if (anc[j] == -1) { anc[j] = k; eparent[j] = k; }
      }
    }
    dbg("etree");
    // postorder (children in ascending order); relabel positions
    std::vector<int> kid_ptr(nb + 1, 0), kid(nb);
    for (int k = 0; k < nb; k++)
      if (eparent[k] >= 0) kid_ptr[eparent[k] + 1]++;
    for (int k = 0; k < nb; k++) kid_ptr[k + 1] += kid_ptr[k];
    {
      std::vector<int> cur(kid_ptr.begin(), kid_ptr.end() - 1);
      for (int k = 0; k < nb; k++)
        if (eparent[k] >= 0) kid[cur[eparent[k]]++] = k;
    }
    std::vector<int> postpos(nb, -1), next_kid(kid_ptr.begin(), kid_ptr.end() - 1), stack;
    int cnt = 0;
    for (int r = 0; r < nb; r++) {
      if (eparent[r] >= 0) continue;
      stack.push_back(r);
      while (!stack.empty()) {
        const int t = stack.back();
        if (next_kid[t] < kid_ptr[t + 1]) stack.push_back(kid[next_kid[t]++]);
        else { postpos[t] = cnt++; stack.pop_back(); }
      }
    }
    std::vector<int> order2(nb), eparent2(nb, -1);
    for (int k = 0; k < nb; k++) {
      order2[postpos[k]] = order[k];
      eparent2[postpos[k]] = eparent[k] >= 0 ? postpos[eparent[k]] : -1;
    }
    order.swap(order2);
    eparent.swap(eparent2);
    for (int k = 0; k < nb; k++) epos[order[k]] = k;
  }
  dbg("postorder");
  // column structures (positions > k, sorted) in one flat pool: struct(k) = later neighbours of the
  // vertex united with the structures of k's etree children minus k itself.  A child always precedes
  // its parent in the postorder, so its structure is already in the pool.
  std::vector<int> cs_ptr(nb + 1, 0), cs;
  {
    // Subtrees of the elimination tree are contiguous in the postorder and independent of each other: the maximal
    // subtrees of at most nb / 16 columns are built side by side on the pool, each into its own pool of entries, the
    // columns above them (the tops of the tree) afterwards; then one flat array.  Which subtree a column belongs to is
    // a property of the tree, so the result does not depend on the thread count.
    std::vector<int> kid_head(nb, -1), kid_next(nb, -1), first_desc(nb);
    for (int k = nb - 1; k >= 0; k--)  // child lists in ascending order
      if (eparent[k] >= 0) { kid_next[k] = kid_head[eparent[k]]; kid_head[eparent[k]] = k; }
    for (int k = 0; k < nb; k++) first_desc[k] = kid_head[k] >= 0 ? first_desc[kid_head[k]] : k;  // subtree of k = [first_desc[k], k]
    const int cap = std::max(64, nb / 16);
    std::vector<int> task_root, owner(nb, -1);  // owner[k] = task that builds column k, -1 = the sequential top
    for (int k = 0; k < nb; k++) {
      const int p = eparent[k];
      const bool fits = k - first_desc[k] + 1 <= cap;
      const bool parent_fits = p >= 0 && p - first_desc[p] + 1 <= cap;
      if (fits && !parent_fits) {
        for (int q = first_desc[k]; q <= k; q++) owner[q] = (int)task_root.size();
        task_root.push_back(k);
      }
    }
    const int ntask = (int)task_root.size();
    std::vector<std::vector<int>> tcs(ntask);            // entries of every task, columns in order
    std::vector<int> col_off(nb + 1, 0), col_len(nb, 0);  // offset of column k inside its owner's pool
    auto build = [&](int k, std::vector<int>& out, std::vector<int>& mark, const std::vector<int>* const* pools) {
      const int v = order[k];
      const size_t s0 = out.size();
      for (int p = nd.xadj[v]; p < nd.xadj[v + 1]; p++) {
        const int q = epos[nd.adj[p]];
        if (q > k && mark[q] != k) { mark[q] = k; out.push_back(q); }
      }
      for (int c = kid_head[k]; c >= 0; c = kid_next[c]) {
        const std::vector<int>& src = *pools[owner[c] + 1];  // may be `out` itself: index, never keep a pointer
        const size_t e0 = (size_t)col_off[c];
        for (int t = 0; t < col_len[c]; t++) {
          const int q = src[e0 + t];
          if (q != k && mark[q] != k) { mark[q] = k; out.push_back(q); }
        }
      }
      std::sort(out.begin() + s0, out.end());
      col_off[k] = (int)s0;
      col_len[k] = (int)(out.size() - s0);
    };
    std::vector<int> top;  // entries of the sequential top columns
    std::vector<const std::vector<int>*> pools(ntask + 1);
    pools[0] = &top;
    for (int t = 0; t < ntask; t++) pools[t + 1] = &tcs[t];
    static std::atomic<unsigned long long> analysis_counter{0};
    const unsigned long long epoch = ++analysis_counter;
    HostPool::get().run(ntask, [&](int t) {
      // one scratch array per pool thread and analysis, not per task (a task marks with its column numbers, which
      // are unique within one analysis)
      thread_local std::vector<int> mark;
      thread_local unsigned long long mark_epoch = 0;
      if (mark_epoch != epoch || (int)mark.size() != nb) { mark.assign(nb, -1); mark_epoch = epoch; }
      tcs[t].reserve(1024);
      for (int k = first_desc[task_root[t]]; k <= task_root[t]; k++) build(k, tcs[t], mark, pools.data());
    });
    {
      std::vector<int> mark(nb, -1);
      for (int k = 0; k < nb; k++)
        if (owner[k] < 0) build(k, top, mark, pools.data());
    }
    for (int k = 0; k < nb; k++) cs_ptr[k + 1] = cs_ptr[k] + col_len[k];
    cs.resize((size_t)cs_ptr[nb]);
    for (int k = 0; k < nb; k++) {
      const int* e = pools[owner[k] + 1]->data() + col_off[k];
      std::copy(e, e + col_len[k], cs.begin() + cs_ptr[k]);
    }
  }
  auto cs_size = [&](int k) { return cs_ptr[k + 1] - cs_ptr[k]; };
  dbg("column structures");
  // supernodes: fundamental (parent[k] == k+1 and struct(k) == {k+1} U struct(k+1)), then relaxed
  // amalgamation of a last child into its parent when the padding it introduces is small
  std::vector<int> snode_first;  // first column of every supernode
  {
    snode_first.push_back(0);
    for (int k = 0; k + 1 < nb; k++) {
      bool merge = eparent[k] == k + 1 && cs_size(k) == cs_size(k + 1) + 1;
      if (!merge && eparent[k] == k + 1) {
        // relaxed: column k+1 starts a supernode whose first column has struct(k+1); padding if k
        // joins = rows of (k+1's front) not in struct(k)
        long sk = 0, sk1 = 0;
        for (int t = cs_ptr[k]; t < cs_ptr[k + 1]; t++) sk += dim[order[cs[t]]];
        for (int t = cs_ptr[k + 1]; t < cs_ptr[k + 2]; t++) sk1 += dim[order[cs[t]]];
        long pad = (sk1 + dim[order[k + 1]]) - sk;  // extra rows carried by the columns merged so far
        long cols = 0;
        for (int c = snode_first.back(); c <= k; c++) cols += dim[order[c]];
        if (pad * cols <= 64 || (pad <= 6 && cols <= 48)) merge = true;
      }
      if (!merge) snode_first.push_back(k + 1);
    }
    if (nb == 0) snode_first.clear();
  }
  // hand the supernode partition to the generic front builder below, in elimination order
  nd.nodes.clear();
  for (size_t sidx = 0; sidx < snode_first.size(); sidx++) {
    int c0 = snode_first[sidx], c1 = sidx + 1 < snode_first.size() ? snode_first[sidx + 1] : nb;
    NdNode node;
    node.verts.assign(order.begin() + c0, order.begin() + c1);
    nd.nodes.push_back(std::move(node));
  }
  int nf = (int)nd.nodes.size();
  std::vector<int> front_of(nb, -1);  // owning node of every block
  for (int f = 0; f < nf; f++)
    for (int v : nd.nodes[f].verts) front_of[v] = f;
  dbg("supernodes");
  // ---- pass 1: update sets of every front and the assembly tree ----
  // The update set of a supernode is the column structure of its LAST column (every earlier column of
  // the chain, and every child subtree, has its structure beyond the chain contained in it), already
  // sorted by elimination position; its first entry names the parent front.
  std::vector<std::vector<int>> U(nf);       // update blocks of every node (sorted by position)
  std::vector<std::vector<int>> akids(nf);   // assembly-tree children
  std::vector<int> aparent(nf, -1);
  for (int f = 0; f < nf; f++) {
    const int last = (f + 1 < nf ? snode_first[f + 1] : nb) - 1;
    std::vector<int>& u = U[f];
    u.reserve(cs_size(last));
    for (int t = cs_ptr[last]; t < cs_ptr[last + 1]; t++) u.push_back(order[cs[t]]);
    if (!u.empty()) {
      const int pf = front_of[u[0]];
      aparent[f] = pf;
      akids[pf].push_back(f);
    }
  }
  dbg("pass1 update sets");
  // ---- latency-driven amalgamation: children on the critical path are absorbed by their parent ----
  // On a GPU a front costs a fixed ~15 us (launch, index chain, zero, scatter, extend-add, write-out)
  // however few pivots it has, and every level of the assembly tree is a dependent launch: the 10-lap
  // trackdrive graph has three levels of fronts with 6 pivots each, the 1-lap graph a chain of a dozen
  // fronts with ~17.  Unlike the chain merge above this also merges across siblings: the pivots of
  // children in different subtrees do not touch each other, so eliminating [child A pivots, child B
  // pivots, parent pivots] inside ONE dense front is the same factorisation with explicit zeros in the
  // A-B block (they stay exactly zero).  The merged front has the parent's update set (struct(child)
  // minus the parent's pivots is contained in it).  Greedy, bottom-up, on a cost model of the numeric
  // kernels (microseconds; calibrated on the ncu launch list of the 10-lap graph,
  // profiles/r01_c2_launches.md): finish(f) = max over children finish(c) + T(f); the children within
  // 10 % of the latest finish are merged into f together whenever that makes f finish earlier.
  // MEASURED AND SWITCHED OFF (symbolic.h): the model's fixed cost per front is too high -- on B200 the
  // merged trees are slower.  Kept behind SLAM_B200_AMALG=t_fix,t_kid,t_panel0,t_panel2,t_area,max_rows
  // (e.g. 14,1.5,0.9,2.0,4.0,150) for re-calibration; tests/test_symbolic.py runs it through the emulation.
  if (amalgamate || getenv("SLAM_B200_AMALG")) {
    double t_fix = 14.0, t_kid = 1.5, t_panel0 = 0.9, t_panel2 = 2.0, t_area = 4.0;
    int max_fs = 150;
    if (const char* e = getenv("SLAM_B200_AMALG")) {
      double v[5];
      int m = 0;
      if (sscanf(e, "%lf,%lf,%lf,%lf,%lf,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &m) == 6) {
        t_fix = v[0]; t_kid = v[1]; t_panel0 = v[2]; t_panel2 = v[3]; t_area = v[4]; max_fs = m;
      }
    }
    auto cost = [&](int sp, int fs, int nk) {
      const double a2 = (fs / 138.0) * (fs / 138.0);
      return t_fix + t_kid * nk + ((sp + 7) / 8) * (t_panel0 + t_panel2 * a2) + t_area * a2;
    };
    std::vector<int> spiv(nf, 0), supd(nf, 0);
    for (int f = 0; f < nf; f++) {
      for (int v : nd.nodes[f].verts) spiv[f] += dim[v];
      for (int w : U[f]) supd[f] += dim[w];
    }
    std::vector<double> fin(nf, 0.0);
    std::vector<char> dead(nf, 0);
    int ndead = 0;
    for (int p = 0; p < nf && max_fs > 0; p++) {  // nodes are in elimination order: children first
      for (;;) {
        std::vector<int>& kids = akids[p];
        double latest = 0.0;
        for (int k : kids) latest = std::max(latest, fin[k]);
        const double now = latest + cost(spiv[p], spiv[p] + supd[p], (int)kids.size());
        if (kids.empty()) { fin[p] = now; break; }
        // candidate: absorb every child within 10 % of the latest finish
        int add = 0, nk = 0;
        double rest = 0.0;
        for (int k : kids) {
          if (fin[k] >= 0.9 * latest) { add += spiv[k]; nk += (int)akids[k].size(); for (int q : akids[k]) rest = std::max(rest, fin[q]); }
          else { nk++; rest = std::max(rest, fin[k]); }
        }
        const int fs2 = spiv[p] + add + supd[p];
        const double then = rest + cost(spiv[p] + add, fs2, nk);
        if (fs2 > max_fs || then >= now - 0.5) { fin[p] = now; break; }
        std::vector<int> keep, front;  // absorbed children's pivots go first, in child order
        for (int k : kids) {
          if (fin[k] >= 0.9 * latest) {
            front.insert(front.end(), nd.nodes[k].verts.begin(), nd.nodes[k].verts.end());
            for (int q : akids[k]) { aparent[q] = p; keep.push_back(q); }
            dead[k] = 1;
            ndead++;
            nd.nodes[k].verts.clear();
            akids[k].clear();
            U[k].clear();
          } else {
            keep.push_back(k);
          }
        }
        std::vector<int>& pv = nd.nodes[p].verts;
        pv.insert(pv.begin(), front.begin(), front.end());
        spiv[p] += add;
        std::sort(keep.begin(), keep.end());
        kids.swap(keep);
      }
    }
    if (ndead) {
      std::vector<int> remap(nf, -1);
      int m = 0;
      for (int f = 0; f < nf; f++)
        if (!dead[f]) remap[f] = m++;
      std::vector<NdNode> nodes2(m);
      std::vector<std::vector<int>> U2(m), kids2(m);
      std::vector<int> par2(m, -1);
      for (int f = 0; f < nf; f++) {
        if (dead[f]) continue;
        const int g = remap[f];
        nodes2[g].verts.swap(nd.nodes[f].verts);
        U2[g].swap(U[f]);
        par2[g] = aparent[f] >= 0 ? remap[aparent[f]] : -1;
        for (int k : akids[f]) kids2[g].push_back(remap[k]);
        std::sort(kids2[g].begin(), kids2[g].end());
      }
      nd.nodes.swap(nodes2);
      U.swap(U2);
      akids.swap(kids2);
      aparent.swap(par2);
      nf = m;
    }
  }
  S.nf = nf;
  std::vector<int> post(nf);
  std::iota(post.begin(), post.end(), 0);
  dbg("amalgamation");
  // ---- pass 2: levels of the assembly tree, level-major renumbering ----
  std::vector<int> level(nf, 0);
  for (int f : post)
    for (int c : akids[f]) level[f] = std::max(level[f], level[c] + 1);
  int nlevels = 0;
  for (int f = 0; f < nf; f++) nlevels = std::max(nlevels, level[f] + 1);
  S.nlevels = nlevels;
  std::vector<int> newid(nf, -1), oldid(nf, -1);
  S.level_ptr.assign(nlevels + 1, 0);
  for (int f = 0; f < nf; f++) S.level_ptr[level[f] + 1]++;
  for (int l = 0; l < nlevels; l++) S.level_ptr[l + 1] += S.level_ptr[l];
  {
    std::vector<int> cur(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int f : post) {  // keeps the postorder inside a level (locality)
      newid[f] = cur[level[f]]++;
      oldid[newid[f]] = f;
    }
  }
  // final positions
  S.pos.assign(nb, -1);
  S.boff.assign(nb, -1);
  S.piv0.assign(nf, 0);
  S.npiv.assign(nf, 0);
  S.nupd.assign(nf, 0);
  S.parent.assign(nf, -1);
  std::vector<int> fof(nb, -1);  // new front id of every block
  {
    int p = 0, off = 0;
    for (int g = 0; g < nf; g++) {
      int f = oldid[g];
      S.piv0[g] = off;
      for (int v : nd.nodes[f].verts) {
        S.pos[v] = p++;
        S.boff[v] = off;
        off += dim[v];
        fof[v] = g;
      }
      S.npiv[g] = off - S.piv0[g];
      S.parent[g] = aparent[f] >= 0 ? newid[aparent[f]] : -1;
    }
    S.n = off;
  }
  // children lists (new ids)
  S.child_ptr.assign(nf + 1, 0);
  for (int g = 0; g < nf; g++)
    if (S.parent[g] >= 0) S.child_ptr[S.parent[g] + 1]++;
  for (int g = 0; g < nf; g++) S.child_ptr[g + 1] += S.child_ptr[g];
  S.children.assign(S.child_ptr[nf], 0);
  {
    std::vector<int> cur(S.child_ptr.begin(), S.child_ptr.end() - 1);
    for (int g = 0; g < nf; g++)
      if (S.parent[g] >= 0) S.children[cur[S.parent[g]]++] = g;
  }
  dbg("pass2 levels/positions");
  // update rows (scalar), sorted by final position
  S.rows_ptr.assign(nf + 1, 0);
  for (int g = 0; g < nf; g++) {
    std::vector<int>& u = U[oldid[g]];
    std::sort(u.begin(), u.end(), [&](int a, int b) { return S.pos[a] < S.pos[b]; });
    int cnt = 0;
    for (int w : u) cnt += dim[w];
    S.nupd[g] = cnt;
    S.rows_ptr[g + 1] = S.rows_ptr[g] + cnt;
  }
  S.upd_rows.assign(S.rows_ptr[nf], 0);
  S.rel.assign(S.rows_ptr[nf], -1);
  for (int g = 0; g < nf; g++) {
    int q = S.rows_ptr[g];
    for (int w : U[oldid[g]])
      for (int k = 0; k < dim[w]; k++) S.upd_rows[q++] = S.boff[w] + k;
  }
  // rel of every child of p / assembly entries of front g: independent per front, each job owns a
  // scratch row map -- contiguous front ranges on the host pool (result independent of the thread count)
  HostPool& pool = HostPool::get();
  const int njobs = std::max(1, std::min(nf, 4 * pool.size()));
  auto job_range = [&](int j, int& f0, int& f1) { f0 = (int)((long)nf * j / njobs); f1 = (int)((long)nf * (j + 1) / njobs); };
  auto fill_pos = [&](std::vector<int>& rp, int g, bool set) {
    const int f = oldid[g];
    int r = 0;
    for (int v : nd.nodes[f].verts) { rp[v] = set ? r : -1; r += dim[v]; }
    for (int w : U[f]) { rp[w] = set ? r : -1; r += dim[w]; }
  };
  pool.run(njobs, [&](int j) {
    int f0, f1;
    job_range(j, f0, f1);
    std::vector<int> rp(nb, -1);
    for (int p = f0; p < f1; p++) {
      if (S.child_ptr[p] == S.child_ptr[p + 1]) continue;
      fill_pos(rp, p, true);
      for (int ci = S.child_ptr[p]; ci < S.child_ptr[p + 1]; ci++) {
        const int g = S.children[ci];
        int q = S.rows_ptr[g];
        for (int w : U[oldid[g]])
          for (int k = 0; k < dim[w]; k++) S.rel[q++] = rp[w] + k;
      }
      fill_pos(rp, p, false);
    }
  });
  dbg("rows+rel");
  // storage offsets, statistics
  S.lptr.assign(nf + 1, 0);
  S.uptr.assign(nf + 1, 0);
  S.nnzL = 0;
  S.flops = 0;
  S.max_front = 0;
  for (int g = 0; g < nf; g++) {
    long s = S.npiv[g], u = S.nupd[g], fs = s + u;
    S.lptr[g + 1] = S.lptr[g] + fs * s;
    S.uptr[g + 1] = S.uptr[g] + u * u;
    S.nnzL += s * (s - 1) / 2 + s * u;
    for (long k = 0; k < s; k++) {
      double m = (double)(fs - k - 1);
      S.flops += m * (m + 1) + m;  // rank-1 update of the trailing lower triangle + scaling
    }
    S.max_front = std::max<int>(S.max_front, (int)fs);
  }
  dbg("offsets/stats");
  ready.fire(true);  // tree, fronts, row maps and storage offsets are final: the caller's launch lists may start
  // ---- assembly entries: every H block lands in the front of its earlier-eliminated vertex ----
  // counting sort by front (diagonal blocks first, then the off-diagonal blocks in input order); the
  // row of the later vertex inside the front is looked up in a per-front scratch map
  S.asm_ptr.assign(nf + 1, 0);
  std::vector<int> off_front(nnb);
  for (int b = 0; b < nb; b++) S.asm_ptr[fof[b] + 1]++;
  for (int k = 0; k < nnb; k++) {
    const int a = off_a[k], b = off_b[k];
    const int g = fof[S.pos[a] < S.pos[b] ? a : b];
    off_front[k] = g;
    S.asm_ptr[g + 1]++;
  }
  for (int g = 0; g < nf; g++) S.asm_ptr[g + 1] += S.asm_ptr[g];
  S.asm_entries.resize(S.asm_ptr[nf]);
  {
    std::vector<int> cur(S.asm_ptr.begin(), S.asm_ptr.end() - 1);
    for (int b = 0; b < nb; b++) {
      const int g = fof[b];
      AsmEntry& e = S.asm_entries[cur[g]++];
      e.hoff = hoff_diag[b];
      e.r = e.c = S.boff[b] - S.piv0[g];
      e.meta = dim[b] | (dim[b] << 8) | (1 << 17);
    }
    // off-diagonal blocks grouped by front, so the row map is filled once per front
    std::vector<int> oc(cur), olist(S.asm_ptr[nf]), obase(cur);  // olist is indexed in entry space
    for (int k = 0; k < nnb; k++) olist[oc[off_front[k]]++] = k;
    pool.run(njobs, [&](int j) {
      int f0, f1;
      job_range(j, f0, f1);
      std::vector<int> rp(nb, -1);
      for (int g = f0; g < f1; g++) {
        const int q0 = obase[g], q1 = S.asm_ptr[g + 1];
        if (q0 == q1) continue;
        fill_pos(rp, g, true);
        for (int q = q0; q < q1; q++) {
          const int k = olist[q];
          const int a = off_a[k], b = off_b[k];
          const bool aEarlier = S.pos[a] < S.pos[b];
          const int e = aEarlier ? a : b, l = aEarlier ? b : a;
          AsmEntry& en = S.asm_entries[q];
          en.hoff = hoff_off[k];
          en.c = S.boff[e] - S.piv0[g];
          en.r = rp[l];
          // stored block is dim[a] x dim[b] (rows a).  We need rows l, columns e.
          const int trans = (l == a) ? 0 : 1;
          en.meta = dim[a] | (dim[b] << 8) | (trans << 16);
        }
        fill_pos(rp, g, false);
      }
    });
  }
  dbg("asm entries");
  S.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
