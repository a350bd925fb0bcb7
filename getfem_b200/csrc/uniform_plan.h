// uniform_plan.h -- host side of the class-uniform tile plan (recompute_uniform.cu); plain C++, no CUDA.
//
// A COLUMN CLASS is the set of column nodes whose descriptors are word-for-word equal.  The descriptor of a column
// node J lists, in CSC order, everything the per-nonzero kernel needs to know about the column except WHERE it is:
//   [pairs] [incidences] [jc[J+1]-jc[J], .., jc[J+Q]-jc[J]]
//   per pair: [keep mask | local contributions << 16] [prel_0 .. prel_{Q-1}]
//             per local contribution: [rank of its element in the column's incidence list << 16 | j*nd + i]
// Members of a class (translated copies of a node in a structured part of a mesh) run in lock step, one per lane,
// through ONE program built here from the descriptor of the class leader; what differs between the lanes -- the CSC
// base jc[J] and the elements of the incidence list -- is per-lane data.
//
// Program of a (class, sub-range of its pairs), 32-bit words:
//   [0] row stride of the tile image (doubles per lane)     [1] pieces (1 .. Q)
//   [2+3b ..] piece b: first entry relative to jc[J], entries, position inside the lane's image row (even)
//   [11] tasks       [12 + g] offset of task g's record (from the program start)
//   task record: [pairs | steps << 8]
//                per pair : [keep mask | piece of component 0 << 16 | of component 1 << 18 | of component 2 << 20]
//                           [offset of the pair's first kept entry of component b inside the row, parity excluded] x 3
//                per step : [rank | code of pair 0 << 16] [code of pair 1 | code of pair 2 << 16]
// A task = up to KGU pairs of the column that are fed by the SAME elements: the geometry row of a step is loaded once
// and serves all of them (register-level operand reuse).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace gf {
namespace uplan {

constexpr int KGU = 3;   // pairs per task
constexpr int HDR = 12;  // header words in front of the task offset table

struct Sub {
  uint32_t prog;    // word offset of the program
  uint32_t ntasks;
  uint32_t weight;  // cost estimate of one tile running this program (arbitrary units)
  uint32_t rowstride;
};

struct ClassPlan {
  int m = 0;  // incidences of a member column
  std::vector<Sub> subs;
};

inline uint32_t row_stride_for(uint32_t rowlen) {  // smallest stride >= rowlen with stride % 16 == 2 (bank spread, 16-byte rows)
  uint32_t s = rowlen <= 2 ? 2 : rowlen;
  while (s % 16 != 2) ++s;
  return s;
}

struct PairD {
  uint32_t mask, cntl;
  uint32_t prel[3];
  const uint32_t *codes;
};

// Appends the programs of one class to `prog`.  row_cap = largest row stride that fits the image buffers,
// task_cap = contributions per task above which a group of pairs is cut.  Returns false with `err` set on failure.
inline bool build_class(const uint32_t *d, size_t dlen, int Q, int nd, uint32_t row_cap, int task_cap,
                        std::vector<uint32_t> &prog, ClassPlan &out, std::string &err) {
  if (dlen < (size_t)(2 + Q)) { err = "short descriptor"; return false; }
  const uint32_t np = d[0];
  out.m = (int)d[1];
  uint32_t coloff[4] = {0, 0, 0, 0};
  for (int b = 0; b < Q; ++b) coloff[b + 1] = d[2 + b];
  std::vector<PairD> pr(np);
  size_t idx = 2 + Q;
  for (uint32_t p = 0; p < np; ++p) {
    if (idx + 1 + Q > dlen) { err = "descriptor overrun"; return false; }
    pr[p].mask = d[idx] & 0xffffu;
    pr[p].cntl = d[idx] >> 16;
    for (int b = 0; b < 3; ++b) pr[p].prel[b] = b < Q ? d[idx + 1 + b] : 0;
    idx += 1 + Q;
    pr[p].codes = d + idx;
    idx += pr[p].cntl;
    if (idx > dlen) { err = "descriptor overrun"; return false; }
    for (uint32_t c = 0; c < pr[p].cntl; ++c)
      if ((pr[p].codes[c] & 0xffffu) >= (uint32_t)(nd * nd) || (pr[p].codes[c] >> 16) >= d[1]) { err = "bad contribution code"; return false; }
  }
  if (idx != dlen) { err = "descriptor length mismatch"; return false; }
  auto start_of = [&](uint32_t p, int b) { return p < np ? coloff[b] + pr[p].prel[b] : coloff[b + 1]; };
  struct Piece { uint32_t goff, len, pbase; };
  auto layout = [&](uint32_t a, uint32_t e, Piece *pc, int &npc) -> uint32_t {  // returns the row length
    if (a == 0 && e == np) {
      npc = 1;
      pc[0] = {0u, coloff[Q], 0u};
      return (coloff[Q] + 2u) & ~1u;
    }
    npc = Q;
    uint32_t at = 0;
    for (int b = 0; b < Q; ++b) {
      pc[b].goff = start_of(a, b);
      pc[b].len = start_of(e, b) - pc[b].goff;
      pc[b].pbase = at;
      at += (pc[b].len + 2u) & ~1u;
    }
    return at;
  };
  uint32_t a = 0;
  while (a < np) {
    Piece pc[3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    int npc = 0;
    uint32_t e = np;
    if (row_stride_for(layout(a, e, pc, npc)) > row_cap) {  // the rest of the column does not fit: longest fitting prefix
      e = a + 1;
      while (e < np && row_stride_for(layout(a, e + 1, pc, npc)) <= row_cap) ++e;
      // even out the remaining ranges so that the last one is not a sliver
      uint32_t rest = np - a, per = e - a, nsub = (rest + per - 1) / per;
      e = a + (rest + nsub - 1) / nsub;
    }
    const uint32_t rowlen = layout(a, e, pc, npc);
    const uint32_t stride = row_stride_for(rowlen);
    if (stride > row_cap) { err = "a single node pair does not fit the tile image"; return false; }
    // ---- groups of pairs with the same element list
    std::map<std::vector<uint16_t>, std::vector<uint32_t>> groups;
    for (uint32_t p = a; p < e; ++p) {
      std::vector<uint16_t> key(pr[p].cntl);
      for (uint32_t c = 0; c < pr[p].cntl; ++c) key[c] = (uint16_t)(pr[p].codes[c] >> 16);
      groups[key].push_back(p);
    }
    struct Task { std::vector<uint32_t> pairs; uint32_t steps, weight; };
    std::vector<Task> tasks;
    for (auto &g : groups) {
      const uint32_t steps = (uint32_t)g.first.size();
      if (steps > 0xffffffu) { err = "too many contributions in one pair"; return false; }
      int cs = steps ? task_cap / (int)steps : KGU;
      cs = std::max(1, std::min(KGU, cs));
      for (size_t q = 0; q < g.second.size(); q += cs) {
        Task t;
        t.steps = steps;
        for (size_t k = q; k < std::min(g.second.size(), q + cs); ++k) t.pairs.push_back(g.second[k]);
        t.weight = 8u + (uint32_t)t.pairs.size() * (steps * 6u + 3u) + steps * 2u;
        tasks.push_back(t);
      }
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const Task &x, const Task &y) { return x.weight > y.weight; });
    if (tasks.size() > 0xffffu) { err = "too many tasks in one tile"; return false; }
    // ---- emit
    Sub sub;
    sub.prog = (uint32_t)prog.size();
    sub.ntasks = (uint32_t)tasks.size();
    sub.rowstride = stride;
    sub.weight = 0;
    const size_t base = prog.size();
    prog.resize(base + HDR + tasks.size(), 0u);
    prog[base + 0] = stride;
    prog[base + 1] = (uint32_t)npc;
    for (int b = 0; b < 3; ++b) {
      prog[base + 2 + 3 * b] = pc[b].goff;
      prog[base + 3 + 3 * b] = b < npc ? pc[b].len : 0u;
      prog[base + 4 + 3 * b] = pc[b].pbase;
    }
    prog[base + 11] = (uint32_t)tasks.size();
    for (size_t g = 0; g < tasks.size(); ++g) {
      const Task &t = tasks[g];
      prog[base + HDR + g] = (uint32_t)(prog.size() - base);
      sub.weight += t.weight;
      prog.push_back((uint32_t)t.pairs.size() | (t.steps << 8));
      for (uint32_t p : t.pairs) {
        uint32_t w0 = pr[p].mask, off[3] = {0, 0, 0};
        for (int b = 0; b < Q; ++b) {
          const int piece = npc == 1 ? 0 : b;
          w0 |= (uint32_t)piece << (16 + 2 * b);
          off[b] = pc[piece].pbase + (coloff[b] + pr[p].prel[b] - pc[piece].goff);
        }
        prog.push_back(w0);
        prog.push_back(off[0]); prog.push_back(off[1]); prog.push_back(off[2]);
      }
      for (uint32_t s = 0; s < t.steps; ++s) {
        uint32_t code[KGU] = {0, 0, 0};
        for (size_t k = 0; k < t.pairs.size(); ++k) code[k] = pr[t.pairs[k]].codes[s] & 0xffffu;
        const uint32_t rank = pr[t.pairs[0]].codes[s] >> 16;
        prog.push_back(rank | (code[0] << 16));
        prog.push_back(code[1] | (code[2] << 16));
      }
    }
    out.subs.push_back(sub);
    a = e;
  }
  return true;
}

}  // namespace uplan
}  // namespace gf
