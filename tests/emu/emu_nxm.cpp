// Host-thread emulation of ONE CTA of the fused NxM optimal-filter kernel (dp_nxm_kernel.cuh).
// Test infrastructure only: checks the table packing, the channel / template bookkeeping and the barrier placement
// without a GPU.
// input: int32 N, n_events, n_chan, n_templ, pretrigger, ac, lo, hi, outside; double fs; templates [n][m][N];
//        csd [n][n][N][2]; traces [n_events][n][N]
// usage: emu_nxm <in.bin> <out.bin> <f32|f64>
#define DP_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../detprocess_b200/csrc/dp_nxm_plan.hpp"

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

template <class T, int R1, int NCH>
static void run_all(const dpnxm::Setup& s, double scale, int subtract_first, int lo, int hi, int outside, const std::vector<double>& traces,
                    int n_events, std::vector<double>& out) {
    using K = DpNxmKernel<T, R1, NCH>;
    using G = Dp2Geom<T, R1>;
    auto dt = dpnxm::build_tables<T, R1>(s, scale);
    DpNxmParams<T> prm{};
    prm.traces = traces.data();
    prm.ev_stride = (long long)s.n * G::N;
    prm.chan_stride = G::N;
    prm.n_events = n_events;
    prm.n_chan = s.n;
    prm.n_templ = s.m;
    prm.tw1 = dt.tw1.data();
    prm.tw2 = dt.tw2.data();
    prm.tw3 = dt.tw3.data();
    prm.twn = dt.twn.data();
    prm.groups = dt.groups.data();
    prm.chunk3 = dt.chunk3.data();
    prm.g = dt.g.data();
    prm.g_self = dt.g_self.data();
    for (int a = 0; a < s.n; ++a) {
        prm.wd[a] = dt.wd[a].data();
        prm.wd_self[a] = dt.wd_self[a].data();
    }
    for (size_t k = 0; k < dt.wo.size(); ++k) {
        prm.wo[k] = dt.wo[k].data();
        prm.wo_self[k] = dt.wo_self[k].data();
    }
    std::memcpy(prm.cmat, dt.cmat, sizeof(prm.cmat));
    std::memcpy(prm.amat, dt.amat, sizeof(prm.amat));
    prm.pretrigger = s.pretrigger;
    prm.lo = lo;
    prm.hi = hi;
    prm.outside = outside;
    const int grid = 2;
    const long long per_cta = K::scratch_v(s.n, s.m);
    std::vector<cx<T>> scratch((size_t)grid * per_cta);
    prm.scratch = scratch.data();
    prm.scratch_per_cta = per_cta;
    prm.out = out.data();
    prm.n_out = 4 + 2 * s.m;
    prm.scale = scale;
    prm.subtract_first = subtract_first;
    for (int b = 0; b < grid; ++b) {
        std::vector<unsigned char> smem(K::SMEM_BYTES + 64);
        unsigned char* sp = smem.data();
        sp += (64 - (reinterpret_cast<uintptr_t>(sp) & 63)) & 63;
        run_cta(G::NT, b, grid, [&] { K::run(prm, sp); });
    }
}

template <class T> static int main_t(const char* in, const char* outp, bool f32) {
    std::ifstream f(in, std::ios::binary);
    int32_t h[9];
    rd(f, h, 9);
    const int N = h[0], n_events = h[1], n = h[2], m = h[3], pre = h[4], ac = h[5], lo = h[6], hi = h[7], outside = h[8];
    double fs;
    rd(f, &fs, 1);
    std::vector<double> templ((size_t)n * m * N), csd((size_t)n * n * N * 2), traces((size_t)n_events * n * N);
    rd(f, templ.data(), templ.size());
    rd(f, csd.data(), csd.size());
    rd(f, traces.data(), traces.size());
    if (!f) { std::fprintf(stderr, "short input\n"); return 2; }
    const dpnxm::Setup s = dpnxm::make_setup(N, fs, n, m, templ.data(), csd.data(), pre, ac != 0);
    double scale = 1.0;
    int subtract_first = 0;
    if (f32) {
        scale = std::exp2(-std::round(std::log2(dpnxm::typical_rms(s, csd.data()))));
        subtract_first = ac ? 1 : 0;
    }
    std::vector<double> out((size_t)n_events * (4 + 2 * m), -1.0);
#define DP_EMU_RUN(R1_, NCH_) run_all<T, R1_, NCH_>(s, scale, subtract_first, lo, hi, outside, traces, n_events, out)
#define DP_EMU_NCH(R1_)                                                              \
    switch (n) {                                                                     \
        case 1: DP_EMU_RUN(R1_, 1); break;                                           \
        case 2: DP_EMU_RUN(R1_, 2); break;                                           \
        case 3: DP_EMU_RUN(R1_, 3); break;                                           \
        case 4: DP_EMU_RUN(R1_, 4); break;                                           \
        default: std::fprintf(stderr, "unsupported n_chan\n"); return 3;             \
    }
    switch (dpplan2::r1_of(N)) {
        case 2: DP_EMU_NCH(2) break;
        case 4: DP_EMU_NCH(4) break;
        case 8: DP_EMU_NCH(8) break;
        default: std::fprintf(stderr, "unsupported N\n"); return 3;
    }
    std::ofstream o(outp, std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: emu_nxm in out f32|f64\n"); return 1; }
    try {
        if (std::string(argv[3]) == "f32") return main_t<f2>(argv[1], argv[2], true);
        return main_t<double>(argv[1], argv[2], false);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 5;
    }
}
