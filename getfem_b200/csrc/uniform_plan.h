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
// Program of a (class, sub-range of its pairs): an array of 8-byte units (two 32-bit words each)
//   unit 0            : [row stride of the tile image (doubles per lane)] [pieces (1 .. Q) | tasks << 8]
//   unit 1 + b, b < 3 : piece b: [first entry relative to jc[J]] [entries | position inside the lane's image row (even) << 16]
//   unit 4 + t        : task t: [first instruction] [one past its last instruction]   (unit indices from the program start)
//   then the instruction stream.  A task is a sequence of GROUPS; a group = up to KGU node pairs of the column that are fed by
//   the SAME elements (the geometry row of a step is loaded once and serves all of them: register-level operand reuse):
//     STEP   x = first step of its group (bit 0) | rank of the element (bits 7-18, so that x & 0x7ff80 is the byte offset of
//                the rank's row of strip positions in the lane data) | code j*nd+i of pair 0 (19-28) | pairs of the group (29-30)
//            y = code of pair 1 (0-9) | code of pair 2 (10-19)
//     FLUSH  x = keep mask (0-8) | piece of component 0, 1, 2 (9-10, 11-12, 13-14) | pair of the group (15-16) |
//                the pair has no local contribution: store zeros (17) | 1 << 31
//            y = offset of the pair's first kept entry of component 0 | 1 << 10 | 2 << 20 inside the row (parity excluded)
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace gf {
namespace uplan {

constexpr int KGU = 3;         // most pairs per group
constexpr int HDR_UNITS = 4;   // units in front of the task table
constexpr uint32_t OP_FLUSH = 1u << 31;

struct Sub {
  uint32_t prog;    // unit offset of the program
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

// Appends the programs of one class to `prog` (8-byte units as pairs of words).  row_cap = largest row stride that fits the
// image buffers, group_cap = contributions above which a group of pairs is cut, ntasks = tasks per tile = warps of the team
// that runs a tile (the groups are packed into them longest-processing-time first; a task may be empty).  Returns false with
// `err` set on failure.
inline bool build_class(const uint32_t *d, size_t dlen, int Q, int nd, uint32_t row_cap, int group_cap, int ntasks, int kg,
                        int max_units, std::vector<uint32_t> &prog, ClassPlan &out, std::string &err) {
  kg = std::max(1, std::min(kg, KGU));  // pairs per group (the kernel variant's accumulator sets)
  if (dlen < (size_t)(2 + Q)) { err = "short descriptor"; return false; }
  const uint32_t np = d[0];
  out.m = (int)d[1];
  if (d[1] > 4095u || nd * nd > 1024) { err = "column valence / element size beyond the instruction format"; return false; }
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
    if (stride > row_cap || stride > 1023u) { err = "a single node pair does not fit the tile image"; return false; }
    for (int b = 0; b < npc; ++b)
      if (pc[b].len > 0xffffu) { err = "column too long for the instruction format"; return false; }
    // ---- groups of pairs with the same element list
    std::map<std::vector<uint16_t>, std::vector<uint32_t>> bylist;
    for (uint32_t p = a; p < e; ++p) {
      std::vector<uint16_t> key(pr[p].cntl);
      for (uint32_t c = 0; c < pr[p].cntl; ++c) key[c] = (uint16_t)(pr[p].codes[c] >> 16);
      bylist[key].push_back(p);
    }
    struct Group { std::vector<uint32_t> pairs; uint32_t steps, weight; };
    std::vector<Group> groups;
    uint32_t wtot = 0, wmax = 0;
    for (auto &g : bylist) {
      const uint32_t steps = (uint32_t)g.first.size();
      int cs = steps ? group_cap / (int)steps : kg;
      cs = std::max(1, std::min(kg, cs));
      for (size_t q = 0; q < g.second.size(); q += cs) {
        Group t;
        t.steps = steps;
        for (size_t k = q; k < std::min(g.second.size(), q + cs); ++k) t.pairs.push_back(g.second[k]);
        t.weight = (uint32_t)t.pairs.size() * (steps * 6u + 4u) + steps * 2u;
        wtot += t.weight;
        wmax = std::max(wmax, t.weight);
        groups.push_back(t);
      }
    }
    // ---- tasks: longest-processing-time packing of the groups into bins of about task_weight
    std::stable_sort(groups.begin(), groups.end(), [](const Group &x, const Group &y) { return x.weight > y.weight; });
    (void)wmax;
    const size_t nbins = (size_t)std::max(1, ntasks);  // one task per warp of the team that runs the tile (some may be empty)
    std::vector<std::vector<size_t>> bins(nbins);
    std::vector<uint32_t> binw(nbins, 0);
    for (size_t g = 0; g < groups.size(); ++g) {
      size_t best = 0;
      for (size_t k = 1; k < nbins; ++k)
        if (binw[k] < binw[best]) best = k;
      bins[best].push_back(g);
      binw[best] += groups[g].weight;
    }
    std::vector<size_t> order(nbins);
    for (size_t k = 0; k < nbins; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return binw[x] > binw[y]; });
    if (nbins > 0xffffu) { err = "too many tasks in one tile"; return false; }
    // ---- emit
    if ((prog.size() / 2) & 1) { prog.push_back(0u); prog.push_back(0u); }  // programs start on 16 bytes
    Sub sub;
    sub.prog = (uint32_t)(prog.size() / 2);
    sub.ntasks = (uint32_t)nbins;
    sub.rowstride = stride;
    sub.weight = *std::max_element(binw.begin(), binw.end()) + 60u;  // a tile lasts as long as its longest task (+ flush)
    const size_t base = prog.size();
    prog.resize(base + 2 * (HDR_UNITS + nbins), 0u);
    prog[base + 0] = stride;
    prog[base + 1] = (uint32_t)npc | ((uint32_t)nbins << 8);
    for (int b = 0; b < 3; ++b) {
      prog[base + 2 * (1 + b)] = pc[b].goff;
      prog[base + 2 * (1 + b) + 1] = (b < npc ? pc[b].len : 0u) | (pc[b].pbase << 16);
    }
    for (size_t t = 0; t < nbins; ++t) {
      if (((prog.size() - base) / 2) & 1) { prog.push_back(0u); prog.push_back(0u); }  // streams start on 16 bytes
      prog[base + 2 * (HDR_UNITS + t)] = (uint32_t)((prog.size() - base) / 2);
      for (size_t gi : bins[order[t]]) {
        const Group &g = groups[gi];
        for (uint32_t s = 0; s < g.steps; ++s) {
          uint32_t code[KGU] = {0, 0, 0};
          for (size_t k = 0; k < g.pairs.size(); ++k) code[k] = pr[g.pairs[k]].codes[s] & 0xffffu;
          const uint32_t rank = pr[g.pairs[0]].codes[s] >> 16;
          prog.push_back((s == 0 ? 1u : 0u) | (rank << 7) | (code[0] << 19) | ((uint32_t)g.pairs.size() << 29));
          prog.push_back(code[1] | (code[2] << 10));
        }
        for (size_t k = 0; k < g.pairs.size(); ++k) {
          const uint32_t p = g.pairs[k];
          uint32_t x = OP_FLUSH | pr[p].mask | ((uint32_t)k << 15) | (g.steps == 0 ? 1u << 17 : 0u), off[3] = {0, 0, 0};
          for (int b = 0; b < Q; ++b) {
            const int piece = npc == 1 ? 0 : b;
            x |= (uint32_t)piece << (9 + 2 * b);
            off[b] = pc[piece].pbase + (coloff[b] + pr[p].prel[b] - pc[piece].goff);
          }
          prog.push_back(x);
          prog.push_back(off[0] | (off[1] << 10) | (off[2] << 20));
        }
      }
      prog[base + 2 * (HDR_UNITS + t) + 1] = (uint32_t)((prog.size() - base) / 2);
      if (prog[base + 2 * (HDR_UNITS + t) + 1] - prog[base + 2 * (HDR_UNITS + t)] > (uint32_t)max_units) {
        err = "a task is longer than the kernel's instruction buffer";
        return false;
      }
    }
    out.subs.push_back(sub);
    a = e;
  }
  return true;
}

}  // namespace uplan
}  // namespace gf
