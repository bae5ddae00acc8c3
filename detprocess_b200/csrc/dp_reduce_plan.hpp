// Host-side plan for the bit-exact window reductions: flattens numpy's pairwise
// summation recursion (pairwise_sum_DOUBLE) into leaves and a level-ordered combine
// tree.  Plain C++; shared by the C-ABI library and the host emulator tests.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>

#include "dp_reduce_kernel.cuh"

namespace dpred {

struct Feat {
    int op, lo, hi;
};

struct Plan {
    int nb_samples = 0;
    double fs = 1.0;
    std::vector<std::vector<Feat>> chan_feats;  // per channel
    // flattened device images
    std::vector<DpRedChan> chans;
    std::vector<DpLeaf> leaves;
    std::vector<DpNode> nodes;
    std::vector<int> level_off;
    std::vector<DpRedFeat> feats;
    int n_out = 0;
    int max_nodes = 1;
};

struct TmpNode {
    int left, right, out, height;
};

// returns node slot holding sum of elements [off, off+n); appends leaves / internal nodes
inline int build_sum(int off, int n, int trapz, int& next_slot, std::vector<DpLeaf>& leaves, std::vector<TmpNode>& inner,
                     int& height_out) {
    if (n <= 128) {
        const int slot = next_slot++;
        leaves.push_back(DpLeaf{off, n, slot, trapz});
        height_out = 0;
        return slot;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    int hl, hr;
    const int l = build_sum(off, n2, trapz, next_slot, leaves, inner, hl);
    const int r = build_sum(off + n2, n - n2, trapz, next_slot, leaves, inner, hr);
    const int slot = next_slot++;
    height_out = std::max(hl, hr) + 1;
    inner.push_back(TmpNode{l, r, slot, height_out});
    return slot;
}

inline void finalize(Plan& p) {
    p.chans.clear();
    p.leaves.clear();
    p.nodes.clear();
    p.level_off.clear();
    p.feats.clear();
    p.n_out = 0;
    p.max_nodes = 1;
    for (const auto& fl : p.chan_feats) {
        DpRedChan ch{};
        ch.leaf_begin = (int)p.leaves.size();
        ch.feat_begin = (int)p.feats.size();
        int next_slot = 0;
        std::vector<TmpNode> inner;
        for (const auto& f : fl) {
            if (f.lo < 0 || f.hi > p.nb_samples || f.hi < f.lo) throw std::invalid_argument("reduce window out of range");
            DpRedFeat df{};
            df.op = f.op;
            df.lo = f.lo;
            df.hi = f.hi;
            df.out = p.n_out++;
            df.n = f.hi - f.lo;
            df.root = 0;
            if (f.op == DP_OP_MEAN || f.op == DP_OP_TRAPZ) {
                const int trapz = f.op == DP_OP_TRAPZ;
                const int n = trapz ? std::max(0, f.hi - f.lo - 1) : (f.hi - f.lo);
                int h;
                df.root = build_sum(f.lo, n, trapz, next_slot, p.leaves, inner, h);
            } else if (f.hi == f.lo) {
                throw std::invalid_argument("zero-size array to reduction operation maximum/minimum which has no identity");
            }
            p.feats.push_back(df);
        }
        ch.leaf_end = (int)p.leaves.size();
        ch.feat_end = (int)p.feats.size();
        // level-order the internal nodes
        std::stable_sort(inner.begin(), inner.end(), [](const TmpNode& a, const TmpNode& b) { return a.height < b.height; });
        ch.level_begin = (int)p.level_off.size();
        size_t i = 0;
        while (i < inner.size()) {
            const int h = inner[i].height;
            p.level_off.push_back((int)p.nodes.size());
            while (i < inner.size() && inner[i].height == h) {
                p.nodes.push_back(DpNode{inner[i].left, inner[i].right, inner[i].out, 0});
                ++i;
            }
        }
        ch.level_end = (int)p.level_off.size();
        p.level_off.push_back((int)p.nodes.size());  // closing offset of this channel's last level
        p.max_nodes = std::max(p.max_nodes, next_slot);
        p.chans.push_back(ch);
    }
    if (p.max_nodes > 4096) throw std::invalid_argument("too many reduction nodes per channel");
}

}  // namespace dpred
