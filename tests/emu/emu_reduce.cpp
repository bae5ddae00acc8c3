// Host-thread emulation of the window-reduction kernel (test infrastructure only).
// usage: emu_reduce <in.bin> <out.bin>
#define DP_HOST_EMU 1
#include <cstdio>
#include <fstream>

#include "../../detprocess_b200/csrc/dp_reduce_plan.hpp"

namespace dpemu {
thread_local Cta* cta = nullptr;
thread_local dp_dim3 tIdx, bIdx, bDim, gDim;
}  // namespace dpemu

template <class F> static void run_cta(int nthreads, int bid, int grid, F&& fn) {
    dpemu::Cta cta(nthreads);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
            dpemu::cta = &cta;
            dpemu::tIdx.x = t;
            dpemu::bIdx.x = bid;
            dpemu::bDim.x = nthreads;
            dpemu::gDim.x = grid;
            fn();
        });
    for (auto& x : th) x.join();
}
template <class V> static void rd(std::ifstream& f, V* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(V) * n); }

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    std::ifstream f(argv[1], std::ios::binary);
    int32_t hdr[3];
    rd(f, hdr, 3);
    const int N = hdr[0], n_events = hdr[1], n_feat = hdr[2];
    double fs;
    rd(f, &fs, 1);
    dpred::Plan plan;
    plan.nb_samples = N;
    plan.fs = fs;
    plan.chan_feats.resize(1);
    for (int i = 0; i < n_feat; ++i) {
        int32_t h[3];
        rd(f, h, 3);
        plan.chan_feats[0].push_back(dpred::Feat{h[0], h[1], h[2]});
    }
    std::vector<double> traces((size_t)n_events * N);
    rd(f, traces.data(), traces.size());
    try {
        dpred::finalize(plan);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 5;
    }
    std::vector<double> out((size_t)n_events * plan.n_out, -1.0);
    DpReduceParams prm{};
    prm.traces = traces.data();
    prm.event_stride = N;
    prm.chan_stride = N;
    prm.nb_samples = N;
    prm.n_rows = n_events;
    prm.n_chan = 1;
    prm.chans = plan.chans.data();
    prm.leaves = plan.leaves.data();
    prm.nodes = plan.nodes.data();
    prm.level_off = plan.level_off.data();
    prm.feats = plan.feats.data();
    prm.out = out.data();
    prm.n_out = plan.n_out;
    prm.fs = fs;
    prm.max_nodes = plan.max_nodes;
    constexpr int NT = 128;
    const int grid = 2;
    for (int b = 0; b < grid; ++b) {
        std::vector<double> nodeval(plan.max_nodes + 8), red(96);
        run_cta(NT, b, grid, [&] { dp_reduce_rows<NT>(prm, nodeval.data(), red.data()); });
    }
    std::ofstream o(argv[2], std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
}
