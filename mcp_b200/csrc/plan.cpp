// plan.cpp — see plan.h.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>

namespace mcpb200 {
namespace {

constexpr int kMaxWindowRows = 256;     // pivot key keeps the slot in 8 bits (kernel_template.cuh)
constexpr int kMaxSensRhs = 16;
constexpr int kSmemBudget = 227 * 1024; // bytes per CTA on sm_100a
constexpr int kAsmChunk = 96;         // terms per chunk of the two-phase assembly (sizes the shared term buffer)
constexpr int kThetaSmemMax = 512;      // θ longer than this is read from global memory in place

bool is_binary(int op) { return op >= MCPB200_OP_ADD && op <= MCPB200_OP_DIV; }
bool is_leaf(int op) { return op >= MCPB200_OP_CONST && op <= MCPB200_OP_THETA; }

struct Entry {  // one non-zero of a sparse block
  int row, col, idx;
};

// ---------------------------------------------------------------------------------------------------
// Reverse Cuthill–McKee on a symmetric adjacency structure, several start nodes per component.
// ---------------------------------------------------------------------------------------------------
std::vector<int> cuthill_mckee_from(const std::vector<std::vector<int>>& adj, int start, std::vector<char>& seen) {
  std::vector<int> order;
  std::queue<int> q;
  q.push(start);
  seen[start] = 1;
  while (!q.empty()) {
    int v = q.front();
    q.pop();
    order.push_back(v);
    std::vector<int> nb;
    for (int u : adj[v])
      if (!seen[u]) {
        seen[u] = 1;
        nb.push_back(u);
      }
    std::sort(nb.begin(), nb.end(), [&](int a, int b) {
      return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b;
    });
    for (int u : nb) q.push(u);
  }
  return order;
}

void bandwidth_of(const std::vector<std::pair<int, int>>& pattern, const std::vector<int>& iperm, int& kl, int& ku) {
  kl = ku = 0;
  for (auto& rc : pattern) {
    int d = iperm[rc.first] - iperm[rc.second];
    if (d > kl) kl = d;
    if (-d > ku) ku = -d;
  }
}

double window_cost(int kl, int ku, int n) {
  double r = std::min(kl + 1, n), c = std::min(kl + ku + 1, n);
  return r * c;
}

// far node of a BFS from `s` (last level, min degree): pseudo-peripheral node heuristic
int far_node(const std::vector<std::vector<int>>& adj, int s, const std::vector<int>& comp_id, int comp) {
  std::vector<int> dist(adj.size(), -1);
  std::queue<int> q;
  q.push(s);
  dist[s] = 0;
  int best = s;
  while (!q.empty()) {
    int v = q.front();
    q.pop();
    if (dist[v] > dist[best] || (dist[v] == dist[best] && adj[v].size() < adj[best].size())) best = v;
    for (int u : adj[v])
      if (dist[u] < 0 && comp_id[u] == comp) {
        dist[u] = dist[v] + 1;
        q.push(u);
      }
  }
  return best;
}

std::vector<int> rcm_ordering(int n, const std::vector<std::pair<int, int>>& pattern) {
  std::vector<std::vector<int>> adj(n);
  {
    std::set<std::pair<int, int>> seen;
    for (auto& rc : pattern) {
      if (rc.first == rc.second) continue;
      int a = std::min(rc.first, rc.second), b = std::max(rc.first, rc.second);
      if (seen.insert({a, b}).second) {
        adj[a].push_back(b);
        adj[b].push_back(a);
      }
    }
  }
  // components
  std::vector<int> comp_id(n, -1);
  int ncomp = 0;
  for (int v = 0; v < n; ++v) {
    if (comp_id[v] >= 0) continue;
    std::queue<int> q;
    q.push(v);
    comp_id[v] = ncomp;
    while (!q.empty()) {
      int w = q.front();
      q.pop();
      for (int u : adj[w])
        if (comp_id[u] < 0) {
          comp_id[u] = ncomp;
          q.push(u);
        }
    }
    ++ncomp;
  }
  std::vector<int> perm;
  perm.reserve(n);
  for (int c = 0; c < ncomp; ++c) {
    std::vector<int> nodes;
    for (int v = 0; v < n; ++v)
      if (comp_id[v] == c) nodes.push_back(v);
    // candidate starts: pseudo-peripheral nodes reached from the min-degree nodes
    std::sort(nodes.begin(), nodes.end(), [&](int a, int b) {
      return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b;
    });
    std::set<int> starts;
    for (size_t i = 0; i < nodes.size() && starts.size() < 12; ++i) {
      int s = nodes[i];
      for (int it = 0; it < 3; ++it) s = far_node(adj, s, comp_id, c);
      starts.insert(s);
      starts.insert(nodes[i]);
    }
    std::vector<int> best_order;
    double best_cost = 1e300;
    std::vector<std::pair<int, int>> sub;
    for (auto& rc : pattern)
      if (comp_id[rc.first] == c && comp_id[rc.second] == c) sub.push_back(rc);
    for (int s : starts) {
      std::vector<char> seen(n, 0);
      std::vector<int> order = cuthill_mckee_from(adj, s, seen);
      std::reverse(order.begin(), order.end());
      std::vector<int> ip(n, 0);
      for (size_t i = 0; i < order.size(); ++i) ip[order[i]] = (int)i;
      int kl, ku;
      bandwidth_of(sub, ip, kl, ku);
      double cost = window_cost(kl, ku, (int)order.size());
      if (cost < best_cost) {
        best_cost = cost;
        best_order = order;
      }
    }
    perm.insert(perm.end(), best_order.begin(), best_order.end());
  }
  return perm;
}

// ---------------------------------------------------------------------------------------------------
// Simulated-annealing refinement of an ordering (start: RCM).  Objective proxy Σ_edges |pos_u − pos_v|^6
// (a smooth stand-in for the bandwidth), swap moves between nearby positions, the true window cost
// (kl+1)(kl+ku+1) is evaluated periodically and the best ordering kept.  Deterministic (fixed seed).
// On the lane-change game this takes the bandwidth from 17 (RCM) to 11–12, which halves the shared-memory
// traffic of every pivot step.
// ---------------------------------------------------------------------------------------------------
std::vector<int> anneal_ordering(int n, const std::vector<std::pair<int, int>>& pattern, std::vector<int> perm) {
  if (n < 8) return perm;
  std::vector<std::vector<int>> adj(n);
  {
    std::set<std::pair<int, int>> seen;
    for (auto& rc : pattern) {
      if (rc.first == rc.second) continue;
      int a = std::min(rc.first, rc.second), b = std::max(rc.first, rc.second);
      if (seen.insert({a, b}).second) {
        adj[a].push_back(b);
        adj[b].push_back(a);
      }
    }
  }
  std::vector<int> pos(n), inv(perm);
  for (int i = 0; i < n; ++i) pos[perm[i]] = i;
  std::vector<double> pw(n + 1);
  for (int d = 0; d <= n; ++d) pw[d] = std::pow((double)d, 6.0);
  auto node_cost = [&](int v, int pv) {
    double c = 0;
    for (int u : adj[v]) c += pw[std::abs(pos[u] - pv)];
    return c;
  };
  int kl, ku;
  bandwidth_of(pattern, pos, kl, ku);
  double best_cost = window_cost(kl, ku, n);
  std::vector<int> best = inv;
  const int bw0 = std::max(kl, ku);
  if (bw0 <= 1) return perm;
  long M = std::max<long>(200000, std::min<long>(4000000, 4000L * n));
  if (const char* e = getenv("MCPB200_ANNEAL_ITERS")) M = std::max<long>(1000, atol(e));   // experiments
  double T = pw[bw0] * 0.5;
  const double alpha = std::exp(std::log(1e-8) / (double)M);
  uint64_t rng = 0x9E3779B97F4A7C15ULL;
  if (const char* e = getenv("MCPB200_ANNEAL_SEED")) rng ^= (uint64_t)atoll(e) * 0xD1B54A32D192ED03ULL;   // experiments
  auto next = [&]() {
    rng ^= rng << 13;
    rng ^= rng >> 7;
    rng ^= rng << 17;
    return rng;
  };
  const int reach = std::max(8, bw0);
  for (long it = 0; it < M; ++it) {
    const int i = (int)(next() % (uint64_t)n);
    int j = i + (int)(next() % (uint64_t)(2 * reach + 1)) - reach;
    j = std::max(0, std::min(n - 1, j));
    if (i != j) {
      const int a = inv[i], b = inv[j];
      const double old_c = node_cost(a, i) + node_cost(b, j);
      pos[a] = j;
      pos[b] = i;
      const double new_c = node_cost(a, j) + node_cost(b, i);
      const double u01 = (double)(next() >> 11) * (1.0 / 9007199254740992.0);
      if (new_c <= old_c || u01 < std::exp(-(new_c - old_c) / T)) {
        inv[i] = b;
        inv[j] = a;
      } else {
        pos[a] = i;
        pos[b] = j;
      }
    }
    T *= alpha;
    if ((it & 4095) == 4095 || it == M - 1) {
      bandwidth_of(pattern, pos, kl, ku);
      const double c = window_cost(kl, ku, n);
      if (c < best_cost) {
        best_cost = c;
        best = inv;
      }
    }
  }
  return best;
}

// ---------------------------------------------------------------------------------------------------
// source emission helpers
// ---------------------------------------------------------------------------------------------------
std::string dlit(double v) {
  char buf[64];
  if (std::isnan(v)) return "__longlong_as_double(0x7ff8000000000000LL)";
  if (std::isinf(v)) return v > 0 ? "__longlong_as_double(0x7ff0000000000000LL)" : "__longlong_as_double(0xfff0000000000000LL)";
  snprintf(buf, sizeof buf, "%.17g", v);
  std::string s(buf);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  return s;
}

// Tables are emitted in the narrowest type that holds them: every index table of the reference's configs fits 16 bits
// and almost every coefficient table is exactly representable in float (±1, small integers) or draws on a handful of
// distinct doubles (a dictionary).  Why: all warps of an SM walk the same ~45 KB of int/double tables once per Newton step
// while only ~28 KB of L1 is left beside the shared memory — packed they are ~19 KB and stay L1-resident (ncu r1g/r2a:
// L1 hit rate 28 %, 22 % of all stall samples were long-scoreboard waits on these loads).
// MCPB200_WIDE_TABLES=1 restores plain int / double tables.
bool wide_tables() {
  const char* e = getenv("MCPB200_WIDE_TABLES");
  return e && atoi(e) != 0;
}

// Hot tables (r2): the index / coefficient tables every warp walks once per Newton step — assembly, condensed right-hand
// side, recovery of δy — can live in STATIC shared memory, copied from their global image when a kernel starts
// (LOAD_HOT_TABLES in load_shared_tables), when that costs no resident instance: the lane-change plan leaves 19 KB of
// the SM's shared memory unused and these tables are 14 KB.  A table named in g_hot.names is then emitted as NAME_G
// (global) plus `__shared__ T NAME[n]`, so the kernels' accesses need no change.
struct HotTables {
  bool on = false;
  std::vector<std::string> names;
  std::ostringstream copy;
  size_t bytes = 0;
  // tables whose name starts with `prefix` (the leaf-index tables of the shape-grouped evaluation, sized only at emission)
  // follow while `room` bytes of the SM's shared memory are left beside the kernels' dynamic blocks
  std::string prefix;
  int64_t room = 0;
  bool has(const std::string& n) const { return on && std::find(names.begin(), names.end(), n) != names.end(); }
};
thread_local HotTables g_hot;

size_t hot_elem_size(const std::string& type) {
  if (type == "short" || type == "unsigned short") return 2;
  if (type == "unsigned char") return 1;
  if (type == "double") return 8;
  return 4;   // int, float
}

template <class T>
void emit_raw_table(std::ostringstream& os, const char* type, const std::string& name, const std::vector<T>& v, bool is_double = false,
                    bool is_float = false) {
  std::string base = name;   // NAME_IDX / NAME_VAL of a dictionary table belong to NAME
  for (const char* suf : {"_IDX", "_VAL"})
    if (base.size() > 4 && base.compare(base.size() - 4, 4, suf) == 0) base = base.substr(0, base.size() - 4);
  bool hot = g_hot.has(base);
  const size_t n = std::max<size_t>(v.size(), 1);
  if (!hot && g_hot.on && g_hot.room > 0 && !g_hot.prefix.empty() && name.rfind(g_hot.prefix, 0) == 0) {
    const int64_t sz = (int64_t)((n * hot_elem_size(type) + 15) & ~size_t(15));
    if (sz <= g_hot.room) {
      hot = true;
      g_hot.room -= sz;
    }
  }
  if (hot) {
    os << "__shared__ " << type << " " << name << "[" << n << "];\n";
    g_hot.copy << "  for (int i_ = threadIdx.x; i_ < " << n << "; i_ += blockDim.x) " << name << "[i_] = " << name << "_G[i_]; \\\n";
    g_hot.bytes += (n * hot_elem_size(type) + 15) & ~size_t(15);
  }
  os << "__device__ const " << type << " " << name << (hot ? "_G" : "") << "[" << std::max<size_t>(v.size(), 1) << "] = {";
  if (v.empty()) os << "0";
  for (size_t i = 0; i < v.size(); ++i) {
    if (i) os << ",";
    if (i % 16 == 15) os << "\n";
    if (is_float) os << dlit((double)v[i]) << "f";
    else if (is_double) os << dlit((double)v[i]);
    else os << (long long)v[i];
  }
  os << "};\n";
}

// `type` "int": narrowed to short when every value fits; "int!" keeps int (tables whose address is taken in the kernels).
// `type` "double": float / dictionary / double, always read through the generated accessor NAME_AT(i).
template <class T>
void emit_table(std::ostringstream& os, const char* type, const char* name, const std::vector<T>& v, bool is_double = false) {
  const std::string ty(type);
  if (!is_double) {
    bool narrow = ty == "int" && !wide_tables();
    for (size_t i = 0; i < v.size() && narrow; ++i) narrow = (long long)v[i] >= -32768 && (long long)v[i] <= 32767;
    emit_raw_table(os, narrow ? "short" : (ty == "int!" ? "int" : type), name, v);
    return;
  }
  const std::string nm(name);
  bool all_float = !wide_tables(), finite = true;
  std::map<unsigned long long, int> dict;
  std::vector<double> vals;
  std::vector<int32_t> idx(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    const double d = (double)v[i];
    if (!std::isfinite(d)) finite = false;
    if (!((double)(float)d == d)) all_float = false;
    unsigned long long bits;
    memcpy(&bits, &d, 8);
    auto it = dict.find(bits);
    if (it == dict.end()) {
      it = dict.emplace(bits, (int)vals.size()).first;
      vals.push_back(d);
    }
    idx[i] = it->second;
  }
  if (all_float && finite) {
    emit_raw_table(os, "float", nm, v, true, true);
    os << "#define " << nm << "_AT(i) ((double)" << nm << "[i])\n";
  } else if (!wide_tables() && vals.size() <= 256 && v.size() > 64) {
    emit_raw_table(os, "unsigned char", nm + "_IDX", idx);
    emit_raw_table(os, "double", nm + "_VAL", vals, true);
    os << "#define " << nm << "_AT(i) (" << nm << "_VAL[" << nm << "_IDX[i]])\n";
  } else {
    emit_raw_table(os, "double", nm, v, true);
    os << "#define " << nm << "_AT(i) (" << nm << "[i])\n";
  }
}

struct Emitter {
  const Plan& P;
  explicit Emitter(const Plan& p) : P(p) {}

  std::string operand(int n) const {
    const int op = P.op[n];
    char buf[64];
    switch (op) {
      case MCPB200_OP_CONST: return "(" + dlit(P.consts[P.a[n]]) + ")";
      case MCPB200_OP_X: snprintf(buf, sizeof buf, "x[%d]", P.a[n]); return buf;
      case MCPB200_OP_Y: snprintf(buf, sizeof buf, "y[%d]", P.a[n]); return buf;
      case MCPB200_OP_THETA: snprintf(buf, sizeof buf, "th[%d]", P.a[n]); return buf;
      default: snprintf(buf, sizeof buf, "v%d", n); return buf;
    }
  }

  // Emits `const double vN = …;` for every non-leaf node the roots need, depth-first per root: each
  // output's expression tree is emitted right before its use, which keeps live ranges (and hence register
  // pressure / spills of the generated device function) short; shared sub-expressions are emitted once.
  void body(std::ostringstream& os, const std::vector<int32_t>& roots) const {
    std::vector<char> done(P.op.size(), 0);
    std::vector<std::pair<int, int>> stack;  // (node, next operand to visit)
    for (int root : roots) {
      if (done[root] || is_leaf(P.op[root])) continue;
      stack.push_back({root, 0});
      while (!stack.empty()) {
        auto& top = stack.back();
        const int n = top.first;
        const int op = P.op[n];
        const int nops = is_binary(op) ? 2 : 1;
        if (top.second < nops) {
          const int child = (top.second == 0) ? P.a[n] : P.b[n];
          ++top.second;
          if (!done[child] && !is_leaf(P.op[child])) stack.push_back({child, 0});
          continue;
        }
        stack.pop_back();
        if (done[n]) continue;
        done[n] = 1;
        statement(os, n);
      }
    }
  }

  // number of not-yet-emitted interior nodes each root would add, processing roots in order
  std::vector<int> incremental_cost(const std::vector<int32_t>& roots) const {
    std::vector<char> done(P.op.size(), 0);
    std::vector<int> cost(roots.size(), 1);
    std::vector<int> stack;
    for (size_t r = 0; r < roots.size(); ++r) {
      stack.assign(1, roots[r]);
      while (!stack.empty()) {
        const int n = stack.back();
        stack.pop_back();
        if (done[n] || is_leaf(P.op[n])) continue;
        done[n] = 1;
        ++cost[r];
        stack.push_back(P.a[n]);
        if (is_binary(P.op[n])) stack.push_back(P.b[n]);
      }
    }
    return cost;
  }

  // Emits `name`_p0 … `name`_p{K-1} (K ≤ 32 balanced, contiguous groups of outputs, each a __noinline__
  // function with bounded register pressure) and the lane dispatcher `name`_par(lane, …): lane i evaluates
  // group i, so the residual / Jacobian evaluation of one instance is spread over the warp.  Sub-expressions
  // shared between groups are recomputed per group.
  // outs[i] = {root node, "target[index]"}.
  // smallest x-leaf (else nx + smallest y-leaf) each node depends on: a locality key — outputs of the same stage of
  // a trajectory problem share sub-expressions, and a group recomputes whatever it shares with other groups
  std::vector<int32_t> locality_keys() const {
    const int32_t big = 1 << 30;
    std::vector<int32_t> key(P.op.size(), big);
    for (size_t n = 0; n < P.op.size(); ++n) {
      const int op = P.op[n];
      if (op == MCPB200_OP_X) key[n] = P.a[n];
      else if (op == MCPB200_OP_Y) key[n] = P.nx + P.a[n];
      else if (is_binary(op)) key[n] = std::min(key[P.a[n]], key[P.b[n]]);
      else if (!is_leaf(op)) key[n] = key[P.a[n]];
    }
    return key;
  }

  // `units` (optional): when the evaluation is large, the part functions go into separately compiled translation
  // units of ~kUnitParts parts each (relocatable device code, linked with nvJitLink by the caller) and `os` only
  // gets their declarations: one translation unit of 30 MB takes NVRTC tens of GB and tens of minutes, the units
  // compile in parallel in a few hundred MB each.
  void partitioned(std::ostringstream& os, const std::string& name, const std::string& params,
                   const std::string& args, std::vector<std::pair<int32_t, std::string>> outs,
                   int max_parts, std::vector<std::string>* units = nullptr, const std::string& unit_prelude = "") const {
    {
      const std::vector<int32_t> key = locality_keys();
      std::stable_sort(outs.begin(), outs.end(), [&](const auto& a, const auto& b) { return key[a.first] < key[b.first]; });
    }
    std::vector<int32_t> roots;
    for (auto& o : outs) roots.push_back(o.first);
    const std::vector<int> cost = incremental_cost(roots);
    long total = 0;
    for (int c : cost) total += c;
    // at most one part per lane while parts stay small; beyond ~kPartNodes tape nodes per part ptxas time explodes
    // (superlinear in function size), so big problems get more parts and every lane loops over several
    constexpr long kPartNodes = 600;
    constexpr int kSplitParts = 96, kUnitParts = 6;   // (cross-unit calls cost: small problems stay in one unit)
    const long by_lanes = std::min<long>(max_parts, total / 24 + 1);
    const int K = (int)std::max<long>(1, std::min<long>((long)outs.size(), std::max<long>(by_lanes, (total + kPartNodes - 1) / kPartNodes)));
    std::vector<size_t> begin(K + 1, outs.size());
    begin[0] = 0;
    {
      long acc = 0;
      int part = 1;
      for (size_t i = 0; i < outs.size() && part < K; ++i) {
        acc += cost[i];
        if (acc * K >= total * part) begin[part++] = i + 1;
      }
    }
    bool split = units && K >= kSplitParts;
    if (const char* e = getenv("MCPB200_SPLIT_COMPILE")) split = units && atoi(e) != 0;
    std::ostringstream unit;
    int in_unit = 0;
    for (int k = 0; k < K; ++k) {
      std::ostringstream& dst = split ? unit : os;
      if (split && in_unit == 0) unit << unit_prelude;
      dst << "__device__ __noinline__ void " << name << "_p" << k << "(" << params << ") {\n";
      std::vector<int32_t> sub(roots.begin() + begin[k], roots.begin() + begin[k + 1]);
      body(dst, sub);
      for (size_t i = begin[k]; i < begin[k + 1]; ++i) dst << "  " << outs[i].second << " = " << operand(outs[i].first) << ";\n";
      dst << "}\n";
      if (split) {
        os << "extern __device__ void " << name << "_p" << k << "(" << params << ");\n";
        if (++in_unit == kUnitParts || k == K - 1) {
          units->push_back(unit.str());
          unit.str("");
          unit.clear();
          in_unit = 0;
        }
      }
    }
    os << "__device__ __forceinline__ void " << name << "_par(int lane, " << params << ") {\n";
    os << "#pragma unroll 1\n  for (int part = lane; part < " << K << "; part += " << max_parts << ")\n  switch (part) {\n";
    for (int k = 0; k < K; ++k) os << "    case " << k << ": " << name << "_p" << k << "(" << args << "); break;\n";
    os << "    default: break;\n  }\n}\n";
  }

  // ---- shape-grouped evaluation ---------------------------------------------------------------------------------
  // Trajectory games repeat a handful of expression shapes over stages and player pairs (the masked game at N = 10
  // has 25 k outputs and 200 k tape nodes, but few distinct per-output expression DAGs once the leaf indices are
  // abstracted).  Outputs whose DAG has the same shape are evaluated by ONE function, lanes in lockstep over the
  // instances of the shape, with the leaf indices / constants that differ read from tables: no lane divergence
  // and code size independent of the number of stages.  (The lane-partitioned functions below run 32 different
  // code paths per warp, i.e. serialised, and their code grows with the problem: 109 MB of SASS for that game.)
  struct OutT {
    int32_t node;
    int arr;       // index into the target array names
    int32_t idx;
  };
  struct ShapeOp {
    int op, a, b;  // operand refs: >= 0 interior (local id), < 0 leaf (-1 - local leaf id); POWI: b = exponent
  };
  struct Shape {
    int arr = 0;
    int root = 0;                        // operand ref of the output value
    std::vector<int> leaf_kind;          // MCPB200_OP_{CONST,X,Y,THETA} per local leaf
    std::vector<ShapeOp> ops;
    std::vector<std::vector<int64_t>> leaf_val;   // [instance][leaf]: index, or (for constants) index into P.consts
    std::vector<int32_t> out_idx;
  };
  static constexpr int kMaxShapeNodes = 96;
  static constexpr int kMinShapeInstances = 6;

  // local DAG of one output; false when it has more than kMaxShapeNodes interior nodes
  bool output_shape(int root, std::string& sig, Shape& sh, std::vector<int64_t>& leaves) const {
    std::map<int, int> interior, leaf;
    sh.ops.clear();
    sh.leaf_kind.clear();
    leaves.clear();
    auto ref = [&](int n) -> int {
      if (is_leaf(P.op[n])) {
        auto it = leaf.find(n);
        if (it == leaf.end()) {
          it = leaf.emplace(n, (int)leaf.size()).first;
          sh.leaf_kind.push_back(P.op[n]);
          leaves.push_back(P.a[n]);
        }
        return -1 - it->second;
      }
      return interior.at(n);
    };
    std::vector<std::pair<int, int>> stack;
    if (!is_leaf(P.op[root])) stack.push_back({root, 0});
    while (!stack.empty()) {
      auto& top = stack.back();
      const int n = top.first;
      const int nops = is_binary(P.op[n]) ? 2 : 1;
      if (top.second < nops) {
        const int child = (top.second == 0) ? P.a[n] : P.b[n];
        ++top.second;
        if (!is_leaf(P.op[child]) && !interior.count(child)) stack.push_back({child, 0});
        continue;
      }
      stack.pop_back();
      if (interior.count(n)) continue;
      if ((int)sh.ops.size() >= kMaxShapeNodes) return false;
      ShapeOp o{P.op[n], ref(P.a[n]), is_binary(P.op[n]) ? ref(P.b[n]) : (P.op[n] == MCPB200_OP_POWI ? P.b[n] : 0)};
      interior.emplace(n, (int)sh.ops.size());
      sh.ops.push_back(o);
    }
    sh.root = ref(root);
    std::ostringstream g;
    for (const ShapeOp& o : sh.ops) g << o.op << ":" << o.a << ":" << o.b << ";";
    g << "|";
    for (int k : sh.leaf_kind) g << k << ",";
    g << "|" << sh.root;
    sig = g.str();
    return true;
  }

  // Emits `name`_par(lane, …): shape-grouped functions for the outputs that repeat, the lane-partitioned functions
  // (`partitioned`) for the rest.
  void evaluation(std::ostringstream& os, const std::string& name, const std::string& params, const std::string& args,
                  const std::vector<OutT>& outs, const std::vector<std::string>& arrays, int max_parts,
                  std::vector<std::string>* units, const std::string& unit_prelude, bool use_shapes) const {
    std::vector<Shape> shapes;
    std::map<std::string, int> by_sig;
    std::vector<int> shape_of(outs.size(), -1);
    if (use_shapes) {
      std::string sig;
      Shape tmp;
      std::vector<int64_t> leaves;
      for (size_t i = 0; i < outs.size(); ++i) {
        if (!output_shape(outs[i].node, sig, tmp, leaves)) continue;
        sig += "|" + std::to_string(outs[i].arr);
        auto it = by_sig.find(sig);
        if (it == by_sig.end()) {
          it = by_sig.emplace(sig, (int)shapes.size()).first;
          shapes.push_back(tmp);
          shapes.back().arr = outs[i].arr;
        }
        Shape& sh = shapes[it->second];
        sh.leaf_val.push_back(leaves);
        sh.out_idx.push_back(outs[i].idx);
        shape_of[i] = it->second;
      }
    }
    std::vector<std::pair<int32_t, std::string>> rest;
    for (size_t i = 0; i < outs.size(); ++i) {
      const int sidx = shape_of[i];
      if (sidx < 0 || (int)shapes[sidx].out_idx.size() < kMinShapeInstances)
        rest.push_back({outs[i].node, arrays[outs[i].arr] + "[" + std::to_string(outs[i].idx) + "]"});
    }
    // small sets of shape functions are inlined into the kernel (lane-change: 2 % faster than calls), big ones stay
    // separate functions to keep the kernel's code size and ptxas time bounded
    size_t shape_ops = 0;
    for (const Shape& sh : shapes)
      if ((int)sh.out_idx.size() >= kMinShapeInstances) shape_ops += sh.ops.size() + sh.leaf_kind.size();
    bool inline_shapes = shape_ops <= 3000;
    if (const char* e = getenv("MCPB200_SHAPE_INLINE")) inline_shapes = atoi(e) != 0;
    std::vector<int> used;
    for (size_t sidx = 0; sidx < shapes.size(); ++sidx) {
      const Shape& sh = shapes[sidx];
      const int cnt = (int)sh.out_idx.size();
      if (cnt < kMinShapeInstances) continue;
      used.push_back((int)sidx);
      const int nl = (int)sh.leaf_kind.size();
      // a leaf whose index (or constant value) is the same in every instance becomes a literal
      std::vector<char> uniform(nl, 1);
      for (int k = 0; k < nl; ++k)
        for (int q = 1; q < cnt && uniform[k]; ++q) {
          if (sh.leaf_kind[k] == MCPB200_OP_CONST) uniform[k] = P.consts[sh.leaf_val[q][k]] == P.consts[sh.leaf_val[0][k]];
          else uniform[k] = sh.leaf_val[q][k] == sh.leaf_val[0][k];
        }
      std::vector<int32_t> itab;   // [varying index leaf][instance], then the output indices
      std::vector<double> ctab;    // [varying constant leaf][instance]
      std::vector<int> irow(nl, -1), crow(nl, -1);
      int ni = 0, nc = 0;
      for (int k = 0; k < nl; ++k) {
        if (uniform[k]) continue;
        if (sh.leaf_kind[k] == MCPB200_OP_CONST) {
          crow[k] = nc++;
          for (int q = 0; q < cnt; ++q) ctab.push_back(P.consts[sh.leaf_val[q][k]]);
        } else {
          irow[k] = ni++;
          for (int q = 0; q < cnt; ++q) itab.push_back((int32_t)sh.leaf_val[q][k]);
        }
      }
      for (int q = 0; q < cnt; ++q) itab.push_back(sh.out_idx[q]);
      const std::string fn = name + "_s" + std::to_string(sidx);
      emit_table(os, "int", (fn + "_I").c_str(), itab);
      if (nc) emit_table(os, "double", (fn + "_C").c_str(), ctab, true);
      os << "__device__ " << (inline_shapes ? "__forceinline__" : "__noinline__") << " void " << fn << "(int lane, " << params << ") {\n";
      os << "#pragma unroll 1\n  for (int q = lane; q < " << cnt << "; q += " << max_parts << ") {\n";
      auto leaf_expr = [&](int k) -> std::string {
        const char* arr = sh.leaf_kind[k] == MCPB200_OP_X ? "x" : sh.leaf_kind[k] == MCPB200_OP_Y ? "y" : "th";
        if (sh.leaf_kind[k] == MCPB200_OP_CONST)
          return uniform[k] ? "(" + dlit(P.consts[sh.leaf_val[0][k]]) + ")" : fn + "_C[" + std::to_string(crow[k] * cnt) + " + q]";
        if (uniform[k]) return std::string(arr) + "[" + std::to_string(sh.leaf_val[0][k]) + "]";
        return std::string(arr) + "[" + fn + "_I[" + std::to_string(irow[k] * cnt) + " + q]]";
      };
      for (int k = 0; k < nl; ++k) os << "    const double l" << k << " = " << leaf_expr(k) << ";\n";
      auto rf = [&](int r) -> std::string { return r < 0 ? "l" + std::to_string(-1 - r) : "n" + std::to_string(r); };
      for (size_t i = 0; i < sh.ops.size(); ++i) {
        const ShapeOp& o = sh.ops[i];
        os << "    const double n" << i << " = ";
        const std::string A = rf(o.a);
        switch (o.op) {
          case MCPB200_OP_ADD: os << A << " + " << rf(o.b); break;
          case MCPB200_OP_SUB: os << A << " - " << rf(o.b); break;
          case MCPB200_OP_MUL: os << A << " * " << rf(o.b); break;
          case MCPB200_OP_DIV: os << A << " / " << rf(o.b); break;
          case MCPB200_OP_NEG: os << "-" << A; break;
          case MCPB200_OP_SQRT: os << "sqrt(" << A << ")"; break;
          case MCPB200_OP_EXP: os << "exp(" << A << ")"; break;
          case MCPB200_OP_LOG: os << "log(" << A << ")"; break;
          case MCPB200_OP_SIN: os << "sin(" << A << ")"; break;
          case MCPB200_OP_COS: os << "cos(" << A << ")"; break;
          case MCPB200_OP_POWI: os << "mcp_powi(" << A << ", " << o.b << ")"; break;
          default: os << "0.0"; break;
        }
        os << ";\n";
      }
      os << "    " << arrays[sh.arr] << "[" << fn << "_I[" << ni * cnt << " + q]] = " << rf(sh.root) << ";\n  }\n}\n";
    }
    const bool have_rest = !rest.empty();
    if (have_rest) partitioned(os, name + "_rest", params, args, rest, max_parts, units, unit_prelude);
    os << "// " << outs.size() << " outputs: " << (outs.size() - rest.size()) << " in " << used.size() << " shape groups, "
       << rest.size() << " lane-partitioned\n";
    os << "__device__ __forceinline__ void " << name << "_par(int lane, " << params << ") {\n";
    for (int sidx : used) os << "  " << name << "_s" << sidx << "(lane, " << args << ");\n";
    if (have_rest) os << "  " << name << "_rest_par(lane, " << args << ");\n";
    os << "}\n";
  }

  void statement(std::ostringstream& os, int n) const {
    const std::string A = operand(P.a[n]);
    os << "  const double v" << n << " = ";
    switch (P.op[n]) {
      case MCPB200_OP_ADD: os << A << " + " << operand(P.b[n]); break;
      case MCPB200_OP_SUB: os << A << " - " << operand(P.b[n]); break;
      case MCPB200_OP_MUL: os << A << " * " << operand(P.b[n]); break;
      case MCPB200_OP_DIV: os << A << " / " << operand(P.b[n]); break;
      case MCPB200_OP_NEG: os << "-" << A; break;
      case MCPB200_OP_SQRT: os << "sqrt(" << A << ")"; break;
      case MCPB200_OP_EXP: os << "exp(" << A << ")"; break;
      case MCPB200_OP_LOG: os << "log(" << A << ")"; break;
      case MCPB200_OP_SIN: os << "sin(" << A << ")"; break;
      case MCPB200_OP_COS: os << "cos(" << A << ")"; break;
      case MCPB200_OP_POWI: os << "mcp_powi(" << A << ", " << P.b[n] << ")"; break;
      default: os << "0.0"; break;
    }
    os << ";\n";
  }
};

Operand classify(const Plan& P, int node, std::map<int, int>& slot_of_node, std::vector<int32_t>& slot_nodes) {
  const int op = P.op[node];
  if (op == MCPB200_OP_CONST) return {P.consts[P.a[node]], -1};
  if (op == MCPB200_OP_THETA) return {1.0, -2 - P.a[node]};
  if (op == MCPB200_OP_NEG && P.op[P.a[node]] == MCPB200_OP_THETA) return {-1.0, -2 - P.a[P.a[node]]};
  if (op == MCPB200_OP_MUL) {
    const int a = P.a[node], b = P.b[node];
    if (P.op[a] == MCPB200_OP_CONST && P.op[b] == MCPB200_OP_THETA) return {P.consts[P.a[a]], -2 - P.a[b]};
    if (P.op[b] == MCPB200_OP_CONST && P.op[a] == MCPB200_OP_THETA) return {P.consts[P.a[b]], -2 - P.a[a]};
  }
  auto it = slot_of_node.find(node);
  if (it == slot_of_node.end()) {
    it = slot_of_node.emplace(node, (int)slot_nodes.size()).first;
    slot_nodes.push_back(node);
  }
  return {1.0, it->second};
}


// Substitute x = 0, y = 0 in the tape and fold constants: the residual's constant part G(0;θ), H(0;θ) of a
// problem that is affine in z.  Returns a small tape (only θ leaves) and, per root, its node in that tape.
struct ZeroTape {
  Plan tape;                     // only op / a / b / consts are used (by Emitter)
  std::vector<int32_t> roots;
};

ZeroTape substitute_zero(const Plan& P, const std::vector<int32_t>& roots) {
  ZeroTape Z;
  Plan& T = Z.tape;
  const size_t n = P.op.size();
  std::vector<char> is_c(n, 0);
  std::vector<double> cval(n, 0.0);
  std::vector<int32_t> node(n, -1);
  std::map<double, int32_t> const_node;
  auto mk_const = [&](double v) -> int32_t {
    auto it = const_node.find(v);
    if (it != const_node.end()) return it->second;
    T.consts.push_back(v);
    T.op.push_back(MCPB200_OP_CONST);
    T.a.push_back((int32_t)T.consts.size() - 1);
    T.b.push_back(-1);
    return const_node[v] = (int32_t)T.op.size() - 1;
  };
  auto mk = [&](int op, int32_t a, int32_t b) -> int32_t {
    T.op.push_back(op);
    T.a.push_back(a);
    T.b.push_back(b);
    return (int32_t)T.op.size() - 1;
  };
  auto as_node = [&](size_t i) -> int32_t { return is_c[i] ? mk_const(cval[i]) : node[i]; };
  std::vector<char> need(n, 0);
  {
    std::vector<int> st(roots.begin(), roots.end());
    while (!st.empty()) {
      int v = st.back();
      st.pop_back();
      if (need[v]) continue;
      need[v] = 1;
      if (is_binary(P.op[v])) { st.push_back(P.a[v]); st.push_back(P.b[v]); }
      else if (!is_leaf(P.op[v])) st.push_back(P.a[v]);
    }
  }
  for (size_t i = 0; i < n; ++i) {
    if (!need[i]) continue;
    const int op = P.op[i];
    const int a = P.a[i], b = P.b[i];
    auto setc = [&](double v) { is_c[i] = 1; cval[i] = v; };
    switch (op) {
      case MCPB200_OP_CONST: setc(P.consts[a]); break;
      case MCPB200_OP_X: case MCPB200_OP_Y: setc(0.0); break;
      case MCPB200_OP_THETA: node[i] = mk(MCPB200_OP_THETA, a, -1); break;
      case MCPB200_OP_ADD:
        if (is_c[a] && is_c[b]) setc(cval[a] + cval[b]);
        else if (is_c[a] && cval[a] == 0.0) node[i] = node[b];
        else if (is_c[b] && cval[b] == 0.0) node[i] = node[a];
        else node[i] = mk(op, as_node(a), as_node(b));
        break;
      case MCPB200_OP_SUB:
        if (is_c[a] && is_c[b]) setc(cval[a] - cval[b]);
        else if (is_c[b] && cval[b] == 0.0) node[i] = node[a];
        else if (is_c[a] && cval[a] == 0.0) node[i] = mk(MCPB200_OP_NEG, node[b], -1);
        else node[i] = mk(op, as_node(a), as_node(b));
        break;
      case MCPB200_OP_MUL:
        if (is_c[a] && is_c[b]) setc(cval[a] * cval[b]);
        else if ((is_c[a] && cval[a] == 0.0) || (is_c[b] && cval[b] == 0.0)) setc(0.0);
        else if (is_c[a] && cval[a] == 1.0) node[i] = node[b];
        else if (is_c[b] && cval[b] == 1.0) node[i] = node[a];
        else node[i] = mk(op, as_node(a), as_node(b));
        break;
      case MCPB200_OP_DIV:
        if (is_c[a] && is_c[b]) setc(cval[a] / cval[b]);
        else if (is_c[a] && cval[a] == 0.0) setc(0.0);
        else if (is_c[b] && cval[b] == 1.0) node[i] = node[a];
        else node[i] = mk(op, as_node(a), as_node(b));
        break;
      case MCPB200_OP_NEG:
        if (is_c[a]) setc(-cval[a]); else node[i] = mk(op, node[a], -1);
        break;
      case MCPB200_OP_POWI:
        if (is_c[a]) setc(std::pow(cval[a], (double)b)); else node[i] = mk(op, node[a], b);
        break;
      default:  // sqrt, exp, log, sin, cos
        if (is_c[a]) {
          const double v = cval[a];
          setc(op == MCPB200_OP_SQRT ? std::sqrt(v) : op == MCPB200_OP_EXP ? std::exp(v) : op == MCPB200_OP_LOG ? std::log(v)
               : op == MCPB200_OP_SIN ? std::sin(v) : std::cos(v));
        } else {
          node[i] = mk(op, node[a], -1);
        }
    }
  }
  for (int32_t r : roots) Z.roots.push_back(as_node(r));
  return Z;
}

// smallest stride ≥ v with stride ≡ 2 (mod 4): rows are 16-byte aligned and lanes striding over rows hit
// distinct 16-byte bank groups, so 128-bit shared-memory accesses are conflict free
int stride_for(int v) {
  int s = (v + 1) & ~1;
  return (s % 4 == 2) ? s : s + 2;
}

}  // namespace

// ===================================================================================================
int build_plan(const mcpb200_problem_desc& d, const std::string& kernel_template, Plan& P) {
  auto fail = [&](int code, const std::string& msg) {
    P.error = msg;
    return code;
  };
  // ---- copy + validate the IR ----------------------------------------------------------------------
  if (d.nx <= 0 || d.ny < 0 || d.ntheta < 0 || d.n_nodes <= 0 || d.jz_nnz < 0)
    return fail(MCPB200_ERR_INVALID_ARGUMENT, "invalid problem dimensions (need nx > 0, ny >= 0, ntheta >= 0)");
  if (!d.op || !d.a || !d.b || !d.gh_nodes || (d.jz_nnz > 0 && (!d.jz_rows || !d.jz_cols || !d.jz_nodes)))
    return fail(MCPB200_ERR_INVALID_ARGUMENT, "null IR array");
  P.nx = d.nx;
  P.ny = d.ny;
  P.nt = d.ntheta;
  P.op.assign(d.op, d.op + d.n_nodes);
  P.a.assign(d.a, d.a + d.n_nodes);
  P.b.assign(d.b, d.b + d.n_nodes);
  P.consts.assign(d.consts, d.consts + std::max(d.n_consts, 0));
  P.gh_nodes.assign(d.gh_nodes, d.gh_nodes + d.nx + d.ny);
  P.jz_rows.assign(d.jz_rows, d.jz_rows + d.jz_nnz);
  P.jz_cols.assign(d.jz_cols, d.jz_cols + d.jz_nnz);
  P.jz_nodes.assign(d.jz_nodes, d.jz_nodes + d.jz_nnz);
  P.has_jt = d.jt_nnz >= 0;
  if (P.has_jt && d.jt_nnz > 0) {
    if (!d.jt_rows || !d.jt_cols || !d.jt_nodes) return fail(MCPB200_ERR_INVALID_ARGUMENT, "null θ-Jacobian array");
    P.jt_rows.assign(d.jt_rows, d.jt_rows + d.jt_nnz);
    P.jt_cols.assign(d.jt_cols, d.jt_cols + d.jt_nnz);
    P.jt_nodes.assign(d.jt_nodes, d.jt_nodes + d.jt_nnz);
  }
  const int nx = P.nx, ny = P.ny, nt = P.nt, nn = d.n_nodes;
  for (int n = 0; n < nn; ++n) {
    const int op = P.op[n], a = P.a[n], b = P.b[n];
    bool ok = true;
    switch (op) {
      case MCPB200_OP_CONST: ok = a >= 0 && a < d.n_consts; break;
      case MCPB200_OP_X: ok = a >= 0 && a < nx; break;
      case MCPB200_OP_Y: ok = a >= 0 && a < ny; break;
      case MCPB200_OP_THETA: ok = a >= 0 && a < nt; break;
      case MCPB200_OP_ADD: case MCPB200_OP_SUB: case MCPB200_OP_MUL: case MCPB200_OP_DIV:
        ok = a >= 0 && a < n && b >= 0 && b < n; break;
      case MCPB200_OP_NEG: case MCPB200_OP_SQRT: case MCPB200_OP_EXP: case MCPB200_OP_LOG:
      case MCPB200_OP_SIN: case MCPB200_OP_COS: case MCPB200_OP_POWI:
        ok = a >= 0 && a < n; break;
      default: ok = false;
    }
    if (!ok) {
      char buf[128];
      snprintf(buf, sizeof buf, "malformed IR node %d (op %d, a %d, b %d): operands must precede users", n, op, a, b);
      return fail(MCPB200_ERR_INVALID_ARGUMENT, buf);
    }
  }
  auto node_ok = [&](int v) { return v >= 0 && v < nn; };
  for (int v : P.gh_nodes) if (!node_ok(v)) return fail(MCPB200_ERR_INVALID_ARGUMENT, "gh_nodes out of range");
  for (size_t k = 0; k < P.jz_nodes.size(); ++k)
    if (!node_ok(P.jz_nodes[k]) || P.jz_rows[k] < 0 || P.jz_rows[k] >= nx + ny || P.jz_cols[k] < 0 || P.jz_cols[k] >= nx + ny)
      return fail(MCPB200_ERR_INVALID_ARGUMENT, "Jacobian entry out of range");
  for (size_t k = 0; k < P.jt_nodes.size(); ++k)
    if (!node_ok(P.jt_nodes[k]) || P.jt_rows[k] < 0 || P.jt_rows[k] >= nx + ny || P.jt_cols[k] < 0 || P.jt_cols[k] >= nt)
      return fail(MCPB200_ERR_INVALID_ARGUMENT, "θ-Jacobian entry out of range");

  // ---- classify Jacobian entries --------------------------------------------------------------------
  {
    std::map<int, int> slots;
    P.jz_opnd.resize(P.jz_nodes.size());
    for (size_t k = 0; k < P.jz_nodes.size(); ++k) {
      P.jz_opnd[k] = classify(P, P.jz_nodes[k], slots, P.jv_nodes);
      if (P.jz_opnd[k].code == -1) ++P.n_const_entries;
    }
    std::map<int, int> tslots;
    P.jt_opnd.resize(P.jt_nodes.size());
    for (size_t k = 0; k < P.jt_nodes.size(); ++k) P.jt_opnd[k] = classify(P, P.jt_nodes[k], tslots, P.jtv_nodes);
  }

  // ---- blocks of the Jacobian -----------------------------------------------------------------------
  std::vector<Entry> Gx, Gy, Hx, Hy;
  for (size_t k = 0; k < P.jz_nodes.size(); ++k) {
    const int r = P.jz_rows[k], c = P.jz_cols[k];
    if (r < nx && c < nx) Gx.push_back({r, c, (int)k});
    else if (r < nx) Gy.push_back({r, c - nx, (int)k});
    else if (c < nx) Hx.push_back({r - nx, c, (int)k});
    else Hy.push_back({r - nx, c - nx, (int)k});
  }
  // ∇_y H ≢ 0 (the reference accepts any H(x, y; θ), src/mcp.jl:27-52,76-80 — none of its own configs has it): mode B,
  // the (nx+ny)-dimensional system with only δs eliminated.  MCPB200_FULL_Y=1 forces it (tests).
  P.full_y = !Hy.empty();
  if (const char* e = getenv("MCPB200_FULL_Y")) P.full_y = P.full_y || atoi(e) != 0;
  const int N = P.full_y ? nx + ny : nx;
  P.N = N;

  // ---- structure of C = G_x + tol·I − G_y D⁻¹ H_x ------------------------------------------------------
  std::vector<std::vector<Entry>> gy_by_k(ny), hx_by_k(ny);
  for (auto& e : Gy) gy_by_k[e.col].push_back(e);
  for (auto& e : Hx) hx_by_k[e.row].push_back(e);
  struct Term {
    double coef;
    int a, b, k;
  };
  std::map<std::pair<int, int>, std::vector<Term>> dest;  // (old row, old col) → terms
  for (int i = 0; i < N; ++i) dest[{i, i}];               // diagonal always present (tol·I)
  for (auto& e : Gx) {
    const Operand& o = P.jz_opnd[e.idx];
    dest[{e.row, e.col}].push_back({o.coef, o.code, -1, -1});
  }
  if (P.full_y) {
    auto put = [&](int r, int c, int idx) {
      const Operand& o = P.jz_opnd[idx];
      dest[{r, c}].push_back({o.coef, o.code, -1, -1});
    };
    for (auto& e : Gy) put(e.row, nx + e.col, e.idx);
    for (auto& e : Hx) put(nx + e.row, e.col, e.idx);
    for (auto& e : Hy) put(nx + e.row, nx + e.col, e.idx);
    for (int k = 0; k < ny; ++k) dest[{nx + k, nx + k}].push_back({1.0, -1, -1, k});   // + s_k/(y_k+tol): 1·[dinv_k·1]
  } else {
    for (int k = 0; k < ny; ++k)
      for (auto& g : gy_by_k[k])
        for (auto& h : hx_by_k[k]) {
          const Operand& og = P.jz_opnd[g.idx];
          const Operand& oh = P.jz_opnd[h.idx];
          dest[{g.row, h.col}].push_back({-og.coef * oh.coef, og.code, oh.code, k});
        }
  }
  std::vector<std::pair<int, int>> pattern;
  pattern.reserve(dest.size());
  for (auto& kv : dest) pattern.push_back(kv.first);

  // ---- ordering --------------------------------------------------------------------------------------
  {
    std::vector<int> ident(N);
    std::iota(ident.begin(), ident.end(), 0);
    int kl0, ku0;
    bandwidth_of(pattern, ident, kl0, ku0);
    std::vector<int> perm = anneal_ordering(N, pattern, rcm_ordering(N, pattern));
    std::vector<int> ip(N);
    for (int i = 0; i < N; ++i) ip[perm[i]] = i;
    int kl1, ku1;
    bandwidth_of(pattern, ip, kl1, ku1);
    if (window_cost(kl1, ku1, N) < window_cost(kl0, ku0, N)) {
      P.perm = perm;
      P.iperm = ip;
      P.kl = kl1;
      P.ku = ku1;
    } else {
      P.perm = ident;
      P.iperm = ident;
      P.kl = kl0;
      P.ku = ku0;
    }
  }
  P.WC = std::min(P.kl + P.ku + 1, N);
  P.R = std::min(P.kl + 1, N);
  if (P.R > kMaxWindowRows) {
    char buf[200];
    snprintf(buf, sizeof buf, "condensed system needs %d window rows (bandwidth kl=%d) > %d supported by the "
             "shared-memory window kernel", P.R, P.kl, kMaxWindowRows);
    return fail(MCPB200_ERR_UNSUPPORTED, buf);
  }
  if (N >= 65536 || P.WC >= 65536) return fail(MCPB200_ERR_UNSUPPORTED, "reduced dimension ≥ 65536");
  // Dense problems: when the whole condensed matrix is resident (no banding to exploit) and the Schur
  // product has far more terms than the matrix has entries, it is accumulated in-kernel as rank-1 updates
  // (kernel_template.cuh, DENSE_SCHUR) instead of term by term.
  {
    size_t n_schur = 0;
    for (auto& kv : dest)
      for (auto& t : kv.second) n_schur += (t.k >= 0);
    P.dense_schur = (!P.full_y && P.R == N && P.WC == N && n_schur > 4 * dest.size()) ? 1 : 0;
    if (const char* e = getenv("MCPB200_DENSE_SCHUR")) P.dense_schur = (atoi(e) != 0) && P.R == N && P.WC == N && !P.full_y;
    P.dense_kernel = (P.dense_schur && N <= 112 && ny <= 128) ? 1 : 0;   // thread mappings of the dense kernel
    if (const char* e = getenv("MCPB200_DENSE_KERNEL")) P.dense_kernel = (atoi(e) != 0) && P.dense_schur && N <= 112 && ny <= 128;
    if (P.dense_schur)
      for (auto& kv : dest) {
        auto& v = kv.second;
        v.erase(std::remove_if(v.begin(), v.end(), [](const Term& t) { return t.k >= 0; }), v.end());
      }
  }
  // Problems of a few unknowns (the README QP: nx = ny = 2): one THREAD per instance, everything in registers, all
  // tables turned into straight-line code (kernel_template.cuh, TINY_KERNEL).  MCPB200_TINY=0 disables it.
  P.tiny_kernel = (!P.full_y && !P.dense_kernel && !P.dense_schur && N <= 6 && ny <= 8 && nt <= 32 && P.op.size() <= 4000) ? 1 : 0;
  if (const char* e = getenv("MCPB200_TINY")) P.tiny_kernel = P.tiny_kernel && atoi(e) != 0;
  P.nrhs_sens = P.has_jt ? std::max(1, std::min(nt, kMaxSensRhs)) : 1;
  if (P.has_jt) {
    // right-hand sides per factorisation pass: as many (≤ 16) as keep one instance within shared memory
    auto sens_bytes = [&](int r) {
      const int64_t fixed = (int64_t)nx + 3 * (int64_t)ny + (int64_t)P.jv_nodes.size() + (int64_t)P.jtv_nodes.size() +
                            (nt <= kThetaSmemMax ? nt : 0) + 64;
      const int64_t win = std::max<int64_t>((int64_t)P.R * (P.WC + r + 4), 1024);
      return 8 * (fixed + (int64_t)r * (ny + N) + win);
    };
    while (P.nrhs_sens > 1 && sens_bytes(P.nrhs_sens) > kSmemBudget - 8192) P.nrhs_sens /= 2;
  }
  P.WS1 = stride_for(P.WC + 1);
  P.WSS = stride_for(P.WC + P.nrhs_sens);

  // ---- assembly tables (new ordering) -------------------------------------------------------------------
  {
    struct D {
      int row, col;
      const std::vector<Term>* terms;
    };
    std::vector<D> ds;
    for (auto& kv : dest) ds.push_back({P.iperm[kv.first.first], P.iperm[kv.first.second], &kv.second});
    std::sort(ds.begin(), ds.end(), [](const D& a, const D& b) { return a.row != b.row ? a.row < b.row : a.col < b.col; });
    P.d_tptr.push_back(0);
    for (auto& dd : ds) {
      P.d_row.push_back(dd.row);
      P.d_col.push_back(dd.col);
      P.d_cpos.push_back(dd.col % P.WC);
      P.d_diag.push_back(dd.row == dd.col);
      double base = 0.0;
      for (auto& t : *dd.terms) {
        if (t.a == -1 && t.k < 0) {  // numeric constant: folded on the host
          base += t.coef;
          continue;
        }
        P.t_coef.push_back(t.coef);
        P.t_a.push_back(t.a);
        P.t_b.push_back(t.b);
        P.t_k.push_back(t.k);
      }
      P.d_tptr.push_back((int)P.t_coef.size());
      P.d_base.push_back(base);
    }
    // rhs rows
    std::vector<std::vector<Entry>> gy_by_row(nx);
    for (auto& e : Gy) gy_by_row[e.row].push_back(e);
    P.r_ptr.push_back(0);
    for (int i = 0; i < N; ++i) {
      const int old = P.perm[i];
      P.r_grow.push_back(old);
      if (!P.full_y)
      for (auto& e : gy_by_row[old]) {
        P.r_coef.push_back(P.jz_opnd[e.idx].coef);
        P.r_code.push_back(P.jz_opnd[e.idx].code);
        P.r_k.push_back(e.col);
      }
      P.r_ptr.push_back((int)P.r_coef.size());
    }
    P.h_ptr.push_back(0);
    for (int k = 0; k < ny; ++k) {
      if (!P.full_y)
      for (auto& e : hx_by_k[k]) {
        P.h_coef.push_back(P.jz_opnd[e.idx].coef);
        P.h_code.push_back(P.jz_opnd[e.idx].code);
        P.h_col.push_back(P.iperm[e.col]);
      }
      P.h_ptr.push_back((int)P.h_coef.size());
    }
    // G_y by column k, rows in the new ordering (dense Schur accumulation)
    P.gk_ptr.push_back(0);
    for (int k = 0; k < ny; ++k) {
      for (auto& e : gy_by_k[k]) {
        P.gk_row.push_back(P.iperm[e.row]);
        P.gk_coef.push_back(P.jz_opnd[e.idx].coef);
        P.gk_code.push_back(P.jz_opnd[e.idx].code);
      }
      P.gk_ptr.push_back((int)P.gk_row.size());
    }
    // θ-Jacobian by column
    std::vector<std::vector<int>> by_q(std::max(nt, 1));
    for (size_t k = 0; k < P.jt_nodes.size(); ++k) by_q[P.jt_cols[k]].push_back((int)k);
    P.q_ptr.push_back(0);
    for (int q = 0; q < nt; ++q) {
      for (int k : by_q[q]) {
        P.q_row.push_back(P.jt_rows[k]);
        P.q_code.push_back(P.jt_opnd[k].code);
        P.q_coef.push_back(P.jt_opnd[k].coef);
      }
      P.q_ptr.push_back((int)P.q_row.size());
    }
  }

  // ---- structure used by the dense kernel v2: G_y = −H_xᵀ (KKT structure), H_x independent of z ---------------
  {
    std::map<std::pair<int, int>, Operand> hx_map;
    bool zconst = true;
    for (auto& e : Hx) {
      hx_map[{e.row, e.col}] = P.jz_opnd[e.idx];
      if (P.jz_opnd[e.idx].code >= 0) zconst = false;
    }
    bool mhxt = Gy.size() == Hx.size();
    for (auto& e : Gy) {
      auto it = hx_map.find({e.col, e.row});
      if (it == hx_map.end() || it->second.code != P.jz_opnd[e.idx].code || it->second.coef != -P.jz_opnd[e.idx].coef) {
        mhxt = false;
        break;
      }
    }
    P.gy_is_mhxt = mhxt;
    P.hx_zconst = zconst;
    P.affine = P.jv_nodes.empty();   // every Jacobian entry is a constant or ±c·θ_i ⇒ G, H are affine in z
  }
  // ---- shared-memory layout and launch configuration ---------------------------------------------------
  P.theta_in_smem = nt <= kThetaSmemMax ? 1 : 0;
  const int njv = (int)P.jv_nodes.size(), njtv = (int)P.jtv_nodes.size();
  std::ostringstream lay;
  auto even = [](int64_t v) { return (v + 1) & ~int64_t(1); };
  int64_t off = 0;
  auto place = [&](const char* name, int64_t n) {
    lay << "#undef " << name << "\n#define " << name << " " << off << "\n";
    off = even(off + n);
  };
  // The "window" region of shared memory is, over one Newton step: storage of G, term buffer of the
  // two-phase assembly, the factorisation window (or, with register-resident rows, just two staging rows),
  // and the cp.async ring of the back substitution.
  // Lanes per instance: a half-warp when the window has at most 16 rows (two instances then share every
  // instruction of the factorisation), else the whole warp.  MCPB200_SUB / MCPB200_REGWIN override (A/B tests).
  P.sub = (P.R <= 8 && P.nrhs_sens <= 16) ? 16 : 32;   // measured: pairing pays for tiny windows (README QP 2x)
  if (const char* e = getenv("MCPB200_SUB")) {
    const int v = atoi(e);
    if (v == 32 || (v == 16 && P.R <= 16 && P.nrhs_sens <= 16)) P.sub = v;
  }
  P.regwin = 1;
  if (const char* e = getenv("MCPB200_REGWIN")) P.regwin = atoi(e) != 0;
  if (P.dense_schur) P.regwin = 0;   // the rank-1 Schur accumulation works on the shared-memory window
  const int64_t nterms_all = (int64_t)P.t_coef.size();
  const int uts = (P.WC + 1) & ~1;   // UT row stride: even ⇒ 16-byte rows for cp.async.cg
  // widest register-window part (entries of one window row held by one lane); MCPB200_REGWIN_PW overrides
  int regwin_pw_max = 48;   // (48: the masked game at N = 4 — 46 entries per lane, 254 registers at 256 threads per SM, no spills — r2: pass 0 −19 %, tail −7 %)
  if (const char* e = getenv("MCPB200_REGWIN_PW")) regwin_pw_max = std::max(2, atoi(e));
  // Geometry of the register-resident window — mirrors BS_NPART / BS_PW / BS_REGWIN of kernel_template.cuh: WR + 1 row
  // slots (one spare, so the entering row is staged a step early), NPART lanes per row, PW positions per lane covering
  // WC + 1 relative columns.
  auto regwin_geom = [&](int& npart, int& pw) -> bool {
    const int spare = 0;   // BS_SPARE of the kernel template (the spare row slot was measured neutral and costs lanes)
    const int rs = P.R + spare;
    npart = (rs <= P.sub) ? ((P.sub / rs) >= 4 ? 4 : ((P.sub / rs) >= 2 ? 2 : 1)) : 1;
    pw = (((P.WC + spare + npart - 1) / npart) + 1) & ~1;
    return P.regwin && rs <= P.sub && pw <= regwin_pw_max;
  };
  // The window kernels keep the G rows in `sol` (the evaluation writes row i to position iperm[i], the condensed
  // right-hand side is then formed in place), so G needs no storage of its own: with the back substitution off the
  // shared-memory ring (r2) the window region shrinks to the assembly's term buffer, and more instances fit an SM.
  const bool g_on_sol = !P.dense_schur && !P.dense_kernel && !P.tiny_kernel;
  auto window_doubles = [&](int ws, int nrhs) -> int64_t {
    int npart, pw;
    const bool regwin = regwin_geom(npart, pw);
    const int es = (pw * npart + nrhs + 3) & ~1;
    if (!regwin) return (int64_t)P.R * ws + even(P.R) + 4;                 // + mailbox of the cooperative sweep (NWIDE)
    int64_t w = 3 * (int64_t)es;                                           // published pivot row + 2 staging rows
    w = std::max<int64_t>(w, std::min<int64_t>(nterms_all, kAsmChunk));     // two-phase assembly buffer (chunked)
    if (!g_on_sol && nx <= 1024) w = std::max<int64_t>(w, nx);             // G alias (plans that do not keep G in `sol`)
    return w;
  };
  const int64_t win_solve = window_doubles(P.WS1, 1);
  int64_t win_sens = window_doubles(P.WSS, P.nrhs_sens);
  // G is consumed (residual norm, condensed rhs) before the window is used, so it shares the window's
  // storage whenever it fits; H[k] is consumed by the very lane/iteration that writes w[k], so H lives in w.
  const bool g_alias = !g_on_sol && win_solve >= nx;
  // `s` is only ever read and written in lane-strided sweeps (residual, recovery of δs, linesearch, update): it can live
  // in a per-instance global block (coalesced, L1/L2-resident) when that buys resident instances (decided below).
  int s_global = 0;
  int64_t solve_state = 0, solve_doubles = 0;
  auto layout_solve = [&]() {
    off = 0;
    place("SOLVE_OFF_X", nx);
    place("SOLVE_OFF_Y", ny);
    if (!s_global) place("SOLVE_OFF_S", ny);
    else lay << "#undef SOLVE_OFF_S\n#define SOLVE_OFF_S 0\n";
    if (!g_alias && !g_on_sol) place("SOLVE_OFF_G", nx);
    place("SOLVE_OFF_JV", njv);
    place("SOLVE_OFF_DINV", ny);
    lay << "#undef SOLVE_OFF_H\n#define SOLVE_OFF_H " << off << "\n";
    place("SOLVE_OFF_W", ny);
    place("SOLVE_OFF_SOL", N);
    if (P.theta_in_smem) place("SOLVE_OFF_TH", nt);
    place("SOLVE_OFF_STAGE", P.dense_schur ? 2 * even(N) : 0);
    solve_state = off;     // everything above is "state"; the window region follows
    if (g_alias || g_on_sol) lay << "#undef SOLVE_OFF_G\n#define SOLVE_OFF_G 0\n";
    lay << "#undef SOLVE_G_IN_WIN\n#define SOLVE_G_IN_WIN " << (g_alias ? 1 : 0) << "\n#undef SOLVE_G_ON_SOL\n#define SOLVE_G_ON_SOL "
        << (g_on_sol ? 1 : 0) << "\n";
    place("SOLVE_OFF_WIN", win_solve);
    solve_doubles = off;
  };
  layout_solve();
  const int64_t stage_n = even(N);
  int64_t sens_state = 0, sens_doubles = 0;
  auto layout_sens = [&]() {
    off = 0;
    place("SENS_OFF_X", nx);
    place("SENS_OFF_Y", ny);
    place("SENS_OFF_S", ny);
    place("SENS_OFF_JV", njv);
    place("SENS_OFF_JTV", njtv);
    place("SENS_OFF_DINV", ny);
    place("SENS_OFF_WQ", (int64_t)P.nrhs_sens * ny);
    place("SENS_OFF_SOL", (int64_t)P.nrhs_sens * N);
    if (P.theta_in_smem) place("SENS_OFF_TH", nt);
    place("SENS_OFF_STAGE", P.dense_schur ? 2 * stage_n : 0);
    sens_state = off;
    place("SENS_OFF_WIN", win_sens);
    sens_doubles = off;
  };
  layout_sens();
  const int64_t nd = (int64_t)P.d_row.size();
  const int64_t shared_table_doubles = even(((int64_t)(N + 1) * 4 + nd * 2 + 7) / 8);
  // Hot tables in static shared memory (see HotTables): sized here with the emitter's own narrowing rules, enabled when
  // the standard layout keeps its number of resident instances with the tables beside it.  MCPB200_HOT_SMEM=0 disables.
  g_hot = HotTables{};
  int64_t hot_reserve = 0;
  {
    auto int_bytes = [&](const std::vector<int32_t>& v) -> int64_t {
      bool narrow = !wide_tables();
      for (size_t i = 0; i < v.size() && narrow; ++i) narrow = v[i] >= -32768 && v[i] <= 32767;
      return (int64_t)((std::max<size_t>(v.size(), 1) * (narrow ? 2 : 4) + 15) & ~size_t(15));
    };
    auto dbl_bytes = [&](const std::vector<double>& v) -> int64_t {
      bool all_float = !wide_tables(), finite = true;
      std::set<unsigned long long> distinct;
      for (double d : v) {
        if (!std::isfinite(d)) finite = false;
        if (!((double)(float)d == d)) all_float = false;
        unsigned long long bits;
        memcpy(&bits, &d, 8);
        distinct.insert(bits);
      }
      const size_t n = std::max<size_t>(v.size(), 1);
      if (all_float && finite) return (int64_t)((n * 4 + 15) & ~size_t(15));
      if (!wide_tables() && distinct.size() <= 256 && v.size() > 64)
        return (int64_t)(((n + 15) & ~size_t(15)) + ((std::max<size_t>(distinct.size(), 1) * 8 + 15) & ~size_t(15)));
      return (int64_t)((n * 8 + 15) & ~size_t(15));
    };
    bool ti16 = !wide_tables();
    for (size_t i = 0; i < P.t_a.size() && ti16; ++i)
      ti16 = std::abs(P.t_a[i]) < 32767 && std::abs(P.t_b[i]) < 32767 && std::abs(P.t_k[i]) < 32767;
    const bool tp16 = !wide_tables() && !P.d_tptr.empty() && P.d_tptr.back() < 32768;
    int64_t est = 0;
    est += (int64_t)((std::max<size_t>(P.d_tptr.size(), 1) * (tp16 ? 2 : 4) + 15) & ~size_t(15));
    est += dbl_bytes(P.d_base) + dbl_bytes(P.t_coef);
    est += (int64_t)((std::max<size_t>(P.t_a.size(), 1) * (ti16 ? 8 : 16) + 15) & ~size_t(15));
    est += int_bytes(P.r_ptr) + int_bytes(P.r_code) + int_bytes(P.r_k) + dbl_bytes(P.r_coef);
    est += int_bytes(P.h_ptr) + int_bytes(P.h_code) + int_bytes(P.h_col) + dbl_bytes(P.h_coef);
    est += int_bytes(P.perm) + int_bytes(P.iperm);
    est += 256 + 2 * 64;   // (+ the chunk tables of the two-phase assembly, a handful of entries)
    int64_t hot_cap = 40 * 1024;   // (static shared memory is limited to 48 KB per kernel)
    if (const char* e = getenv("MCPB200_HOT_CAP")) hot_cap = atoll(e);
    bool want = !P.dense_kernel && !P.dense_schur && !P.tiny_kernel && !P.full_y && est <= hot_cap;
    if (const char* e = getenv("MCPB200_HOT_SMEM")) want = want && atoi(e) != 0;
    if (want) {
      // same number of instances per CTA with and without the reserve (whole quads of warps above 16, as below)?
      auto inst = [&](int64_t reserve) {
        int64_t w = (kSmemBudget - shared_table_doubles * 8 - reserve) / (solve_doubles * 8);
        w = std::min<int64_t>(w, 768 / P.sub);
        if (P.sub == 16) w &= ~int64_t(1);
        if (P.sub == 32 && w > 16) w = (w / 4) * 4;
        return w;
      };
      // … or, for plans that will keep only the window in shared memory (large-state mode, decided below with the
      // same counts), the same number of windows up to that mode's cap
      auto wins = [&](int64_t reserve, int64_t doubles) {
        return std::max<int64_t>(0, std::min<int64_t>((kSmemBudget - shared_table_doubles * 8 - reserve) / (doubles * 8), 768 / P.sub));
      };
      int64_t cap_ls = 8;
      if (const char* e = getenv("MCPB200_LS_WARPS")) cap_ls = std::max(1, atoi(e));
      const int64_t ls_doubles = even(win_solve) + even(N);
      const bool would_ls = inst(0) < 1 || (inst(0) < 4 && std::min(cap_ls, wins(0, even(win_solve))) >= 2 * inst(0));
      const bool std_ok = inst(0) >= 1 && inst(est) == inst(0);
      const bool ls_ok = wins(est, ls_doubles) >= 1 && std::min(cap_ls, wins(est, ls_doubles)) == std::min(cap_ls, wins(0, ls_doubles));
      if (would_ls ? ls_ok : std_ok) {
        g_hot.on = true;
        g_hot.names = {"D_TP", "D_BASE", "T_COEF", "T_I", "R_PTR", "R_CODE", "R_K", "R_COEF", "H_PTR", "H_CODE", "H_COL", "H_COEF",
                       "PERM", "IPERM", "CH_D", "CH_T"};
        hot_reserve = est;
      }
    }
  }
  auto warps_for = [&](int64_t doubles) {  // instances per CTA
    int64_t w = (kSmemBudget - shared_table_doubles * 8 - hot_reserve) / (doubles * 8);
    w = std::min<int64_t>(w, 768 / P.sub);   // ≤ 768 threads per CTA keeps ≥ 85 registers per thread
    if (const char* e = getenv("MCPB200_MAX_WARPS")) w = std::min<int64_t>(w, std::max(1, atoi(e)));   // tuning: fewer instances, more L1
    if (P.sub == 16) w &= ~int64_t(1);       // whole warps
    return (int)std::max<int64_t>(0, w);
  };
  // Large-state mode: when one instance's vectors do not fit shared memory next to the window (e.g. the masked
  // game at N = 10: nx = 3000, ny = 3630), the iterate / residual / step vectors live in an L2-resident global
  // block per instance and only the factorisation window stays in shared memory.  MCPB200_LARGE_STATE=1 forces it.
  P.large_state = 0;
  int ls_cap = 8;   // state traffic goes through L1/L2: a few warps per SM are enough to cover it
  if (const char* e = getenv("MCPB200_LS_WARPS")) ls_cap = std::max(1, atoi(e));
  if (!P.dense_kernel) {
    const int w_std = std::min(warps_for(solve_doubles), P.has_jt ? warps_for(sens_doubles) : 1 << 30);
    const int w_ls = std::min(ls_cap, std::min(warps_for(even(win_solve)), P.has_jt ? warps_for(even(win_sens)) : 1 << 30));
    // also when the vectors crowd the window out of shared memory: the masked game at N = 4 fits 2 instances per
    // SM with its state in shared memory and 8 without (measured: solves +27 %, pullbacks 4.2x)
    if (w_std < 1 || (w_std < 4 && w_ls >= 2 * w_std)) P.large_state = 1;
    if (const char* e = getenv("MCPB200_LARGE_STATE")) P.large_state = atoi(e) != 0;
  }
  if (P.large_state && P.has_jt && !P.dense_kernel) {
    // with the state out of shared memory the right-hand sides per factorisation pass are limited by the window alone
    // (the count above was sized for vectors in shared memory: 1 for the masked games, i.e. one factorisation per
    // column of ∇F_θ for a full Jacobian)
    int r = std::max(1, std::min(nt, kMaxSensRhs));
    const int want = std::max(1, std::min(ls_cap, warps_for(even(win_sens))));
    while (r > P.nrhs_sens && warps_for(even(window_doubles(stride_for(P.WC + r), r))) < want) r /= 2;
    if (r > P.nrhs_sens) {
      P.nrhs_sens = r;
      P.WSS = stride_for(P.WC + r);
      win_sens = window_doubles(P.WSS, r);
      layout_sens();
    }
  }
  // Resident warps come in fours (one register file per scheduler: 17 warps cost the registers of 20), so above 16 only
  // multiples of 4 are used; `s` moves to the global block when that reaches the next multiple.  MCPB200_S_GLOBAL overrides.
  auto quad = [](int w) { return w > 16 ? (w / 4) * 4 : w; };
  if (!P.large_state && !P.dense_kernel && !P.tiny_kernel && !P.dense_schur && P.sub == 32 && ny > 0) {
    int npart, pw;
    if (regwin_geom(npart, pw)) {
      // Measured on the lane-change game (bench, converged solves/s): 20 instances at 96 registers 219.8–227.1 k, 16
      // instances at 128 registers 230.6–237.3 k — the pivot loop spills at 96, and the SM is bound by its shared-memory
      // data pipe, not by the number of resident warps.  So the move is opt-in: MCPB200_S_GLOBAL=1.
      (void)quad(warps_for(solve_doubles - even(ny)));
      s_global = 0;
      if (const char* e = getenv("MCPB200_S_GLOBAL")) s_global = atoi(e) != 0;
      if (s_global) {
        layout_solve();
        P.state_doubles_solve = even(ny) + 2;
      }
    }
  }
  int sol_in_smem = 0;
  // Window in global memory: when not even ONE instance's factorisation window fits shared memory (a dense problem of a
  // few hundred unknowns, a band of a few hundred rows) the window moves to the instance's L2-resident global block as
  // well — the same code on a generic pointer, an order of magnitude slower per pivot step than the shared-memory
  // window, but the problem is solved instead of refused.  MCPB200_WIN_GLOBAL=1 forces it (tests).
  int win_global = 0;
  if (!P.dense_kernel && !P.tiny_kernel) {
    const bool nofit = warps_for(even(win_solve)) < 1 || (P.has_jt && warps_for(even(win_sens)) < 1);
    if (nofit) win_global = 1;
    if (const char* e = getenv("MCPB200_WIN_GLOBAL")) win_global = (atoi(e) != 0) || nofit;
    if (win_global) P.large_state = 1;
  }
  if (P.large_state && win_global) {
    solve_doubles = 2;   // nothing of an instance lives in shared memory (the per-CTA tables still do)
    sens_doubles = 2;
    P.state_doubles_solve = even(solve_state) + even(win_solve) + 2;
    P.state_doubles_sens = even(sens_state) + even(win_sens) + 2;
    ls_cap = std::min(ls_cap, 4);
  } else if (P.large_state) {
    solve_doubles = even(win_solve);
    // δx / the right-hand side is read and written once per pivot step and once per back-substitution step: keep it
    // in shared memory next to the window when that does not cost an instance (masked game N = 4: 9.6 KB)
    const int w0 = std::min(ls_cap, warps_for(solve_doubles));
    if (!P.dense_kernel && w0 >= 1 && std::min(ls_cap, warps_for(solve_doubles + even(N))) >= w0) {
      sol_in_smem = 1;
      lay << "#define SOLVE_SOL_SMEM_OFF " << solve_doubles << "\n";
      solve_doubles += even(N);
    }
    if (const char* e = getenv("MCPB200_SOL_SMEM")) {
      if (!atoi(e) && sol_in_smem) {
        sol_in_smem = 0;
        solve_doubles -= even(N);
      }
    }
    sens_doubles = even(win_sens);
    P.state_doubles_solve = even(solve_state) + 2;
    P.state_doubles_sens = even(sens_state) + 2;
  }
  P.ipc_solve = warps_for(solve_doubles);
  if (!P.large_state && !P.dense_kernel && !P.tiny_kernel && P.sub == 32) P.ipc_solve = quad(P.ipc_solve);
  {
    // a register window wider than 40 entries per lane needs ≈ 250 registers per thread: at most 256 threads per CTA
    int npart, pw;
    if (!P.dense_kernel && !P.tiny_kernel && regwin_geom(npart, pw) && pw > 40) P.ipc_solve = std::min(P.ipc_solve, 256 / P.sub);
  }
  P.ipc_sens = P.has_jt ? warps_for(sens_doubles) : 1;
  if (P.large_state) {
    P.ipc_solve = std::min(P.ipc_solve, ls_cap);
    P.ipc_sens = std::min(P.ipc_sens, ls_cap);
  }
  // Big shared-memory windows (the masked games): one warp per instance leaves the SM at 3–8 warps and every pivot
  // step is a long dependent sweep; NWIDE warps then share the sweep of one instance (column batches round-robin),
  // the instance's other phases stay on its leader warp.  MCPB200_NWIDE overrides.
  P.nwide = 1;
  if (!P.dense_kernel && P.sub == 32 && P.ipc_solve >= 1 && P.theta_in_smem) {
    int npart, pw;
    const bool regwin_used = regwin_geom(npart, pw);
    const int np = (P.WC + 1 + 1) / 2, pb = std::min(np, 9), nbatch = (np + pb - 1) / pb;
    if (!regwin_used && !P.dense_schur) {
      P.nwide = std::max(1, std::min({4, 16 / P.ipc_solve, nbatch}));
      if (const char* e = getenv("MCPB200_NWIDE")) P.nwide = std::max(1, std::min({atoi(e), nbatch, 32 / P.ipc_solve}));
      if (win_global) P.nwide = 1;   // (the helpers' mailbox sits behind a shared-memory window)
      // the helpers of instance slot k meet at named barrier 1 + k: ids 1 … 15 exist (0 is __syncthreads)
      if (P.nwide > 1 && P.ipc_solve > 15) P.nwide = 1;
    }
  }
  if (P.ipc_solve < 1) {
    char buf[200];
    snprintf(buf, sizeof buf, "per-instance working set (%lld bytes) exceeds the %d-byte shared memory of one SM",
             (long long)(solve_doubles * 8), kSmemBudget);
    return fail(MCPB200_ERR_UNSUPPORTED, buf);
  }
  if (P.has_jt && P.ipc_sens < 1) {
    // the solve fits, the sensitivity solve does not: build without it (callers get MCPB200_ERR_NO_SENSITIVITIES)
    P.has_jt = false;
    P.ipc_sens = 1;
  }
  P.smem_solve = (shared_table_doubles + solve_doubles * P.ipc_solve) * 8;
  // Adjoint-mode pullback (kernel_template.cuh, mcp_adj_kernel): the forward layout with ONE right-hand side, so
  // far more instances fit an SM than in the forward kernel (lane-change: 4 → 14).  MCPB200_ADJOINT=0 disables it.
  P.has_adjoint = (P.has_jt && !P.full_y && !win_global && !P.dense_schur && !P.dense_kernel && P.kl == P.ku) ? 1 : 0;   // (mode B: forward solves only)
  if (const char* e = getenv("MCPB200_ADJOINT")) P.has_adjoint = P.has_adjoint && atoi(e) != 0;
  int64_t adj_doubles = 0;
  if (P.has_adjoint) {
    off = 0;
    place("ADJ_OFF_X", nx);
    place("ADJ_OFF_Y", ny);
    place("ADJ_OFF_S", ny);
    place("ADJ_OFF_JV", njv);
    place("ADJ_OFF_JTV", njtv);
    place("ADJ_OFF_DINV", ny);
    place("ADJ_OFF_WQ", ny);
    place("ADJ_OFF_SOL", N);
    if (P.theta_in_smem) place("ADJ_OFF_TH", nt);
    place("ADJ_OFF_STAGE", 0);
    const int64_t adj_state = off;
    place("ADJ_OFF_WIN", win_solve);
    adj_doubles = off;
    if (P.large_state) {
      adj_doubles = even(win_solve);
      P.state_doubles_adj = even(adj_state) + 2;
    }
    P.ipc_adj = warps_for(adj_doubles);
    if (P.large_state) P.ipc_adj = std::min(P.ipc_adj, ls_cap);
    if (P.ipc_adj < 1) P.has_adjoint = 0;
    P.smem_adj = (shared_table_doubles + adj_doubles * std::max(P.ipc_adj, 1)) * 8;
  }
  if (P.dense_kernel) {
    const int dtr = (N + 15) / 16, dnp = 16 * dtr;
    off = 0;
    place("DENSE_OFF_X", nx);
    place("DENSE_OFF_Y", ny);
    place("DENSE_OFF_S", ny);
    place("DENSE_OFF_G", nx);
    place("DENSE_OFF_W", ny);
    place("DENSE_OFF_DINV", ny);
    place("DENSE_OFF_SOL", N);
    place("DENSE_OFF_JV", njv);
    if (P.theta_in_smem) place("DENSE_OFF_TH", nt);
    place("DENSE_OFF_WIN", (int64_t)N * P.WS1);
    place("DENSE_OFF_STG", 2 * 8 * (int64_t)dnp);
    place("DENSE_OFF_PART", 256);
    place("DENSE_OFF_RED", 16);
    place("DENSE_OFF_RD", N);
    place("DENSE_OFF_ORD", (N + 1) / 2 + 1);
    if (off * 8 > kSmemBudget) {
      P.dense_kernel = 0;   // does not fit one CTA: stay on the general kernel
    } else {
      P.smem_solve = off * 8;
      P.ipc_solve = 1;
      P.dense_ctas_per_sm = (int)std::max<int64_t>(1, std::min<int64_t>(2, (kSmemBudget + 1024) / (off * 8 + 1024)));
    }
    // v2: H_x cached in shared memory for the whole solve, every mat-vec and the Schur complement from it
    bool want_v2 = P.dense_kernel && P.gy_is_mhxt && P.hx_zconst && P.affine;
    if (const char* e = getenv("MCPB200_DENSE_KERNEL")) want_v2 = want_v2 && atoi(e) >= 2;
    if (want_v2) {
      const int64_t hcs = (N + 1) | 1;   // odd row stride of the cached H_x
      const int64_t off_v1 = off;
      place("DENSE_OFF_HC", (int64_t)ny * hcs);
      place("DENSE_OFF_XT", N);
      place("DENSE_OFF_G0", nx + ny);
      if (off * 8 <= kSmemBudget) {
        P.dense_kernel = 2;
        P.smem_solve = off * 8;
        P.dense_ctas_per_sm = (int)std::max<int64_t>(1, std::min<int64_t>(2, (kSmemBudget + 1024) / (off * 8 + 1024)));
        lay << "#define DENSE_HCS " << hcs << "\n";
      } else {
        off = off_v1;
      }
    }
    // v3: the matrix lives in register tiles of a 512-thread CTA (32 lanes × ≤4 row chunks, 16 warps × ≤7 columns);
    // shared memory holds H_x, Uᵀ for the back substitution and the multipliers of the current column
    bool want_v3 = P.dense_kernel == 2 && N + 1 <= 112 && ny <= 128;
    if (const char* e = getenv("MCPB200_DENSE_KERNEL")) want_v3 = want_v3 && atoi(e) >= 3;
    if (want_v3) {
      // Uᵀ row stride ≡ 2 (mod 4) and > N: the symmetric path reads and writes column PAIRS (two adjacent steps) as
      // 128-bit accesses, conflict-free across the lanes of a quarter-warp, broadcast for the pivot-row entries
      int64_t utld = N + 1;
      while (utld % 4 != 2) ++utld;
      const int64_t hcs = (N + 1) | 1, kls = (ny + 7) & ~int64_t(7);
      const int64_t off_v2 = off;
      off = 0;
      place("DENSE_OFF_X", nx);
      place("DENSE_OFF_Y", ny);
      place("DENSE_OFF_S", ny);
      place("DENSE_OFF_G", nx);
      place("DENSE_OFF_W", ny);
      place("DENSE_OFF_DINV", ny);
      place("DENSE_OFF_SOL", N);
      place("DENSE_OFF_JV", njv);
      if (P.theta_in_smem) place("DENSE_OFF_TH", nt);
      place("DENSE_OFF_RED", 16);
      place("DENSE_OFF_RD", N);
      place("DENSE_OFF_HC", (int64_t)ny * hcs);
      place("DENSE_OFF_XT", N);
      place("DENSE_OFF_G0", nx + ny);
      place("DENSE_OFF_UT", std::max<int64_t>((int64_t)(N + 1) * utld, 16 * 128));
      place("DENSE_OFF_MBUF", 2 * 128);
      place("DENSE_OFF_KL", (int64_t)N * kls / 8);    // per column of H_x: the constraints with a non-zero entry (bytes)
      place("DENSE_OFF_KCNT", (N + 1) / 2);           // … and how many (ints)
      if (off * 8 <= kSmemBudget) {
        P.dense_kernel = 3;
        P.dense_threads = 512;
        P.smem_solve = off * 8;
        P.dense_ctas_per_sm = 1;
        lay << "#define DENSE_UTLD " << utld << "\n#define DENSE_KLS " << kls << "\n";
      } else {
        off = off_v2;
      }
    }
  }
  P.smem_sens = (shared_table_doubles + sens_doubles * P.ipc_sens) * 8;
  const int64_t cval_doubles = even(nd) + 2;
  P.scratch_doubles_solve = cval_doubles + even((int64_t)N * uts) + 2;
  if (P.dense_kernel == 3) P.scratch_doubles_solve = std::max<int64_t>(P.scratch_doubles_solve, (int64_t)N * 128);
  P.scratch_doubles_sens = P.scratch_doubles_solve;
  // algorithmic flops of one Newton step's KKT solve: banded LU with partial pivoting +
  // forward/back substitution, dense-in-band count; for dense plans ⅔n³ + 2n² plus the Schur product
  // 2·Σ_k nnz(G_y[:,k])·nnz(H_x[k,:])
  {
    const double kl = P.kl, kuu = std::min(P.kl + P.ku, N - 1);
    P.flops_band = 2.0 * N * kl * kuu + 2.0 * N * kl + 2.0 * N * kuu;
    if (P.dense_schur) {
      double schur = 0.0;
      for (int k = 0; k < ny; ++k) schur += 2.0 * (double)gy_by_k[k].size() * (double)hx_by_k[k].size();
      P.flops_band = 2.0 / 3.0 * N * (double)N * N + 2.0 * N * (double)N + schur;
    }
  }

  // ---- source ----------------------------------------------------------------------------------------------
  std::ostringstream os;
  os << "// generated by libmcpb200 (plan.cpp) — do not edit\n";
  if (const char* e = getenv("MCPB200_DEFS")) {   // kernel tuning experiments: "NAME=VALUE,NAME=VALUE"
    std::string d(e), item;
    std::istringstream is(d);
    while (std::getline(is, item, ',')) {
      const size_t eq = item.find('=');
      if (eq != std::string::npos) os << "#define " << item.substr(0, eq) << " " << item.substr(eq + 1) << "\n";
    }
  }
  os << "#define NX " << nx << "\n#define NY " << ny << "\n#define NT " << nt << "\n#define NRED " << N << "\n";
  os << "#define KL " << P.kl << "\n#define KU " << P.ku << "\n#define WC " << P.WC << "\n#define WR " << P.R << "\n";
  os << "#define WS1 " << P.WS1 << "\n#define WSS " << P.WSS << "\n#define NRHS_SENS " << P.nrhs_sens << "\n";
  os << "#define NJV " << njv << "\n#define NJTV " << njtv << "\n#define ND " << P.d_row.size() << "\n";
  os << "#define THETA_IN_SMEM " << P.theta_in_smem << "\n#define HAS_JT " << (P.has_jt ? 1 : 0) << "\n";
  os << "#define SUB " << P.sub << "\n#define SOLVE_INST " << P.ipc_solve << "\n#define SENS_INST " << P.ipc_sens << "\n";
  os << "#define SOLVE_SMEM_DOUBLES " << solve_doubles << "\n#define SENS_SMEM_DOUBLES " << sens_doubles << "\n";
  os << "#define SOLVE_SCRATCH " << P.scratch_doubles_solve << "\n#define SENS_SCRATCH " << P.scratch_doubles_sens << "\n";
  {
    const int64_t win1 = win_solve, wins = win_sens, nterms = (int64_t)P.t_coef.size();
    // the window's storage doubles as the term buffer of the two-phase assembly and as the cp.async ring
    // two-phase assembly in chunks: consecutive dests whose terms fit the shared term buffer
    const int64_t buf = std::min(win1, wins);
    std::vector<int32_t> ch_d, ch_t;   // chunk c covers dests [ch_d[c], ch_d[c+1]) and terms [ch_t[c], ch_t[c+1])
    bool two_phase = nterms > 0;
    ch_d.push_back(0);
    ch_t.push_back(0);
    for (size_t d0 = 0; d0 < P.d_row.size() && two_phase;) {
      size_t d1 = d0;
      while (d1 < P.d_row.size() && P.d_tptr[d1 + 1] - P.d_tptr[d0] <= buf) ++d1;
      if (d1 == d0) {   // a single dest with more terms than the buffer: fall back to the one-phase loop
        two_phase = false;
        break;
      }
      ch_d.push_back((int32_t)d1);
      ch_t.push_back(P.d_tptr[d1]);
      d0 = d1;
    }
    if (!two_phase) {
      ch_d.assign(1, 0);
      ch_t.assign(1, 0);
    }
    os << "#define NTERMS " << nterms << "\n#define ASM_TWO_PHASE " << (two_phase ? 1 : 0) << "\n#define ASM_NCHUNK "
       << (ch_d.size() - 1) << "\n";
    emit_table(os, "int", "CH_D", ch_d);
    emit_table(os, "int", "CH_T", ch_t);
    os << "#define UTS " << uts << "\n#define REGWIN " << P.regwin << "\n";
    os << "#define RING_D " << std::max<int64_t>(2, std::min<int64_t>(8, std::min(win1, wins) / uts)) << "\n";
  }
  os << "#define DENSE_KERNEL " << P.dense_kernel << "\n#define LARGE_STATE " << P.large_state << "\n";
  os << "#define FULL_Y " << P.full_y << "\n#define S_GLOBAL " << s_global << "\n";
  os << "#define TINY_KERNEL " << P.tiny_kernel << "\n";
  os << "#define WIN_GLOBAL " << win_global << "\n";
  os << "#define NWIDE " << P.nwide << "\n#define SOL_IN_SMEM " << sol_in_smem << "\n#define REGWIN_PW_MAX " << regwin_pw_max << "\n";
  os << "#define SOLVE_STATE_DOUBLES " << P.state_doubles_solve << "\n#define SENS_STATE_DOUBLES " << P.state_doubles_sens << "\n";
  os << "#define DENSE_SCHUR " << P.dense_schur << "\n#define STAGE_N " << stage_n << "\n";
  os << "#define CVAL_DOUBLES " << cval_doubles << "\n#define SHARED_TABLE_DOUBLES " << shared_table_doubles << "\n";
  os << lay.str();
  {
    std::vector<int32_t> tp(P.d_tptr.size()), rowptr(N + 1, 0);
    // D_TP: first term of each dest; the sign bit marks a diagonal dest (gets tol).  16-bit when the term count allows.
    const bool tp16 = !wide_tables() && !P.d_tptr.empty() && P.d_tptr.back() < 32768;
    os << "#define D_TP_MASK " << (tp16 ? "0x7fff" : "0x7fffffff") << "\n";
    for (size_t i = 0; i < P.d_tptr.size(); ++i)
      tp[i] = tp16 ? (int32_t)(int16_t)(uint16_t)(P.d_tptr[i] | ((i < P.d_diag.size() && P.d_diag[i]) ? 0x8000 : 0))
                   : (P.d_tptr[i] | ((i < P.d_diag.size() && P.d_diag[i]) ? (int32_t)0x80000000 : 0));
    for (size_t i = 0; i < P.d_row.size(); ++i) rowptr[P.d_row[i] + 1]++;   // dests are sorted by row
    for (int i = 0; i < N; ++i) rowptr[i + 1] += rowptr[i];
    emit_table(os, "int", "D_ROWPTR", rowptr);
    if (P.dense_kernel) emit_table(os, "int", "D_ROW", P.d_row);
    emit_table(os, "int", "D_CPOS", P.d_cpos);
    emit_table(os, tp16 ? "int" : "int!", "D_TP", tp);
    emit_table(os, "double", "D_BASE", P.d_base, true);
    // Adjoint sensitivities (θ̄ from z̄ with ONE solve of Cᵀ instead of nθ solves of C): the same non-zeros in
    // column-major order — first dest of each column, window position of its row, index into Cval.
    const bool adjoint = P.has_adjoint != 0;
    os << "#define HAS_ADJOINT " << (adjoint ? 1 : 0) << "\n";
    os << "#define ADJ_INST " << P.ipc_adj << "\n#define ADJ_SMEM_DOUBLES " << adj_doubles << "\n#define ADJ_STATE_DOUBLES "
       << P.state_doubles_adj << "\n";
    if (adjoint) {
      std::vector<int32_t> ord(P.d_row.size());
      for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int32_t)i;
      std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {
        return P.d_col[a] != P.d_col[b] ? P.d_col[a] < P.d_col[b] : P.d_row[a] < P.d_row[b];
      });
      std::vector<int32_t> tptr(N + 1, 0), tcpos(ord.size()), tsrc(ord.size());
      for (size_t e = 0; e < ord.size(); ++e) {
        tptr[P.d_col[ord[e]] + 1]++;
        tcpos[e] = P.d_row[ord[e]] % P.WC;
        tsrc[e] = ord[e];
      }
      for (int i = 0; i < N; ++i) tptr[i + 1] += tptr[i];
      emit_table(os, "int", "DT_ROWPTR", tptr);
      emit_table(os, "int", "DT_CPOS", tcpos);
      emit_table(os, "int!", "DT_SRC", tsrc);
    }
  }
  emit_table(os, "double", "T_COEF", P.t_coef, true);
  {
    bool ti16 = !wide_tables();
    for (size_t i = 0; i < P.t_a.size() && ti16; ++i)
      ti16 = std::abs(P.t_a[i]) < 32767 && std::abs(P.t_b[i]) < 32767 && std::abs(P.t_k[i]) < 32767;
    os << "typedef " << (ti16 ? "short4" : "int4") << " TI_T;\n";
    const bool ti_hot = g_hot.has("T_I");
    if (ti_hot) {
      const size_t n = std::max<size_t>(P.t_a.size(), 1);
      os << "__shared__ TI_T T_I[" << n << "];\n";
      g_hot.copy << "  for (int i_ = threadIdx.x; i_ < " << n << "; i_ += blockDim.x) T_I[i_] = T_I_G[i_]; \\\n";
      g_hot.bytes += (n * (ti16 ? 8 : 16) + 15) & ~size_t(15);
    }
    os << "__device__ const TI_T " << (ti_hot ? "T_I_G" : "T_I") << "[" << std::max<size_t>(P.t_a.size(), 1) << "] = {";
    if (P.t_a.empty()) os << "{0,0,0,0}";
    for (size_t i = 0; i < P.t_a.size(); ++i) {
      if (i) os << ",";
      if (i % 8 == 7) os << "\n";
      os << "{" << P.t_a[i] << "," << P.t_b[i] << "," << P.t_k[i] << ",0}";
    }
    os << "};\n";
  }
  emit_table(os, "int", "R_GROW", P.r_grow);
  emit_table(os, "int", "R_PTR", P.r_ptr);
  emit_table(os, "int", "R_CODE", P.r_code);
  emit_table(os, "int", "R_K", P.r_k);
  emit_table(os, "double", "R_COEF", P.r_coef, true);
  emit_table(os, "int", "H_PTR", P.h_ptr);
  emit_table(os, "int", "H_CODE", P.h_code);
  emit_table(os, "int", "H_COL", P.h_col);
  emit_table(os, "double", "H_COEF", P.h_coef, true);
  if (P.dense_schur) {
    emit_table(os, "int", "GK_PTR", P.gk_ptr);
    emit_table(os, "int", "GK_ROW", P.gk_row);
    emit_table(os, "int", "GK_CODE", P.gk_code);
    emit_table(os, "double", "GK_COEF", P.gk_coef, true);
  }
  emit_table(os, "int", "PERM", P.perm);
  emit_table(os, "int", "IPERM", P.iperm);
  if (P.has_jt) {
    emit_table(os, "int", "Q_PTR", P.q_ptr);
    emit_table(os, "int", "Q_ROW", P.q_row);
    emit_table(os, "int", "Q_CODE", P.q_code);
    emit_table(os, "double", "Q_COEF", P.q_coef, true);
  }
  if (g_hot.on && (int64_t)g_hot.bytes > hot_reserve) return fail(MCPB200_ERR_INTERNAL, "hot-table shared memory under-estimated");
  const int64_t max_dynamic = std::max<int64_t>({(int64_t)P.smem_solve, P.has_jt ? (int64_t)P.smem_sens : 0, P.has_adjoint ? (int64_t)P.smem_adj : 0});
  if (g_hot.on) {   // what is left goes to the evaluation's leaf-index tables, in emission order
    g_hot.prefix = "mcp_eval_newton_s";
    g_hot.room = std::min<int64_t>(kSmemBudget - max_dynamic - (int64_t)g_hot.bytes - 128, 46 * 1024 - (int64_t)g_hot.bytes);
    if (const char* e = getenv("MCPB200_HOT_EVAL")) if (atoi(e) == 0) g_hot.room = 0;
  }
  const std::string powi_src =
      "__device__ __forceinline__ double mcp_powi(double a, int n) {\n"
      "  double r = 1.0; bool neg = n < 0; if (neg) n = -n;\n"
      "  while (n) { if (n & 1) r *= a; a *= a; n >>= 1; }\n"
      "  return neg ? 1.0 / r : r;\n}\n";
  os << powi_src;
  const std::string unit_prelude = "// generated by libmcpb200 (plan.cpp): evaluation parts, compiled separately — do not edit\n" + powi_src;
  P.units.clear();
  Emitter E(P);
  // residual rows [G; H] and the computed Jacobian entries, evaluated together so sub-expressions are shared
  os << "// G, H (src/mcp.jl:76-80 minus the structural slack rows) and the z/θ-dependent entries of ∇F_z\n";
  bool use_shapes = true;   // MCPB200_SHAPES=0: lane-partitioned functions only
  if (const char* e = getenv("MCPB200_SHAPES")) use_shapes = atoi(e) != 0;
  if (P.dense_kernel < 2) {
    std::vector<Emitter::OutT> outs;
    for (int i = 0; i < nx; ++i) outs.push_back({P.gh_nodes[i], 0, g_on_sol ? P.iperm[i] : i});   // G row i → sol[iperm[i]]
    for (int i = 0; i < ny; ++i) outs.push_back({P.gh_nodes[nx + i], 1, i});
    for (int i = 0; i < njv; ++i) outs.push_back({P.jv_nodes[i], 2, i});
    E.evaluation(os, "mcp_eval_newton",
                 P.large_state ? "const double* x, const double* y, const double* th, double* g, double* h, double* jv"
                               : "const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ th, "
                                 "double* __restrict__ g, double* __restrict__ h, double* __restrict__ jv",
                 "x, y, th, g, h, jv", outs, {"g", "h", "jv"}, P.dense_kernel ? 256 : P.sub * P.nwide, &P.units, unit_prelude, use_shapes);
  }
  if (P.dense_kernel >= 2) {
    os << "// G(0;θ), H(0;θ): the constant part of the (affine in z) residual\n";
    ZeroTape Z = substitute_zero(P, P.gh_nodes);
    Emitter EZ(Z.tape);
    std::vector<std::pair<int32_t, std::string>> outs;
    for (int i = 0; i < nx + ny; ++i) outs.push_back({Z.roots[i], "gh0[" + std::to_string(i) + "]"});
    EZ.partitioned(os, "mcp_eval_const",
                   "const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ th, "
                   "double* __restrict__ gh0",
                   "x, y, th, gh0", outs, 256);
  }
  if (P.has_jt) {
    os << "// computed entries of ∇F_z and ∇F_θ at the solution (src/AutoDiff.jl:27-37)\n";
    std::vector<Emitter::OutT> outs;
    for (int i = 0; i < njv; ++i) outs.push_back({P.jv_nodes[i], 0, i});
    for (int i = 0; i < njtv; ++i) outs.push_back({P.jtv_nodes[i], 1, i});
    if (outs.empty()) {
      os << "__device__ __forceinline__ void mcp_eval_sens_par(int, const double*, const double*, const double*, double*, double*) {}\n";
    } else {
      E.evaluation(os, "mcp_eval_sens",
                   P.large_state ? "const double* x, const double* y, const double* th, double* jv, double* jtv"
                                 : "const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ th, "
                                   "double* __restrict__ jv, double* __restrict__ jtv",
                   "x, y, th, jv, jtv", outs, {"jv", "jtv"}, P.sub, &P.units, unit_prelude, use_shapes);
    }
  }
  g_hot.room = 0;
  if (g_hot.on && max_dynamic + (int64_t)g_hot.bytes > kSmemBudget) return fail(MCPB200_ERR_INTERNAL, "hot tables exceed the shared memory of one SM");
  os << "#define HOT_SMEM " << (g_hot.on ? 1 : 0) << "\n#define LOAD_HOT_TABLES() do { \\\n" << g_hot.copy.str() << "} while (0)\n";
  P.hot_smem_bytes = g_hot.on ? (int)g_hot.bytes : 0;
  if (P.tiny_kernel) {
    auto val = [&](int code) -> std::string {
      if (code == -1) return "1.0";
      if (code >= 0) return "jv[" + std::to_string(code) + "]";
      return "th[" + std::to_string(-2 - code) + "]";
    };
    const std::string A_X = "const double (&x)[NX]", A_Y = "const double (&y)[NY]", A_TH = "const double (&th)[NT > 0 ? NT : 1]",
                      A_JV = "const double (&jv)[NJV > 0 ? NJV : 1]";
    os << "// ---- thread-per-instance kernel: the tables above as straight-line code on register arrays ----\n";
    os << "__device__ __forceinline__ void tiny_eval(" << A_X << ", " << A_Y << ", " << A_TH
       << ", double (&g)[NX], double (&h)[NY], double (&jv)[NJV > 0 ? NJV : 1]) {\n";
    {
      std::vector<int32_t> roots;
      for (int i = 0; i < nx + ny; ++i) roots.push_back(P.gh_nodes[i]);
      for (int i = 0; i < njv; ++i) roots.push_back(P.jv_nodes[i]);
      E.body(os, roots);
      for (int i = 0; i < nx; ++i) os << "  g[" << i << "] = " << E.operand(P.gh_nodes[i]) << ";\n";
      for (int i = 0; i < ny; ++i) os << "  h[" << i << "] = " << E.operand(P.gh_nodes[nx + i]) << ";\n";
      for (int i = 0; i < njv; ++i) os << "  jv[" << i << "] = " << E.operand(P.jv_nodes[i]) << ";\n";
    }
    os << "}\n";
    // C (new ordering) with the right-hand side in column NRED
    os << "__device__ __forceinline__ void tiny_assemble(" << A_JV << ", " << A_TH
       << ", const double (&dinv)[NY], const double tol, double (&C)[NRED][NRED + 1]) {\n";
    os << "#pragma unroll\n  for (int i = 0; i < NRED; ++i)\n#pragma unroll\n    for (int j = 0; j < NRED; ++j) C[i][j] = 0.0;\n";
    for (size_t d = 0; d < P.d_row.size(); ++d) {
      os << "  C[" << P.d_row[d] << "][" << P.d_col[d] << "] = " << dlit(P.d_base[d]);
      if (P.d_diag[d]) os << " + tol";
      for (int t = P.d_tptr[d]; t < P.d_tptr[d + 1]; ++t) {
        os << " + " << dlit(P.t_coef[t]) << " * " << val(P.t_a[t]);
        if (P.t_k[t] >= 0) os << " * dinv[" << P.t_k[t] << "] * " << val(P.t_b[t]);
      }
      os << ";\n";
    }
    os << "}\n";
    os << "__device__ __forceinline__ void tiny_rhs(const double (&g)[NX], " << A_JV << ", " << A_TH
       << ", const double (&w)[NY], double (&C)[NRED][NRED + 1]) {\n";
    for (int i = 0; i < N; ++i) {
      os << "  C[" << i << "][NRED] = -g[" << P.r_grow[i] << "]";
      for (int e = P.r_ptr[i]; e < P.r_ptr[i + 1]; ++e)
        os << " - " << dlit(P.r_coef[e]) << " * " << val(P.r_code[e]) << " * w[" << P.r_k[e] << "]";
      os << ";\n";
    }
    os << "}\n";
    os << "__device__ __forceinline__ void tiny_hx(" << A_JV << ", " << A_TH
       << ", const double (&v)[NRED], double (&hx)[NY]) {\n";
    for (int k = 0; k < ny; ++k) {
      os << "  hx[" << k << "] = 0.0";
      for (int e = P.h_ptr[k]; e < P.h_ptr[k + 1]; ++e)
        os << " + " << dlit(P.h_coef[e]) << " * " << val(P.h_code[e]) << " * v[" << P.h_col[e] << "]";
      os << ";\n";
    }
    os << "}\n";
    os << "__device__ __forceinline__ void tiny_update_x(double (&x)[NX], const double (&v)[NRED], const double a) {\n";
    for (int c = 0; c < N; ++c) os << "  x[" << P.perm[c] << "] += a * v[" << c << "];\n";
    os << "}\n";
  }
  os << kernel_template;
  P.source = os.str();
  return MCPB200_OK;
}

}  // namespace mcpb200
