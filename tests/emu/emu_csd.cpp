// Host-thread emulation of the noise-CSD accumulation kernel (dp_csd_kernel.cuh), two CTAs, followed by the fold over
// CTAs and the thread-order -> natural-bin map on the host.  Test infrastructure only.
// input: int32 N, n_events, n_chan, f32 (0/1); double fs, typical_rms; uint8 mask[n_events]; traces [n_events][n][N]
// output: double sums[n*n][N/2+1]; double count
// usage: emu_csd <in.bin> <out.bin>
#define DP_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../detprocess_b200/csrc/dp_csd_kernel.cuh"
#include "../../detprocess_b200/csrc/dp_plan2.hpp"

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
static void run_all(double fs, double scale, int subtract_first, const std::vector<double>& traces, const std::vector<unsigned char>& mask,
                    int n_events, std::vector<double>& sums, double& count) {
    using K = DpCsdKernel<T, R1, NCH>;
    using G = Dp2Geom<T, R1>;
    std::vector<dpplan::Channel> none;
    auto dt = dpplan2::build_tables2<T, R1>(fs, none, 0.0, 1.0);
    const std::vector<int> loc = dpplan2::partial_slot_of_bin<G>();
    const int grid = 2;
    const long long ppc = K::PARTIAL * K::NCP;
    std::vector<double> partial((size_t)grid * ppc, 0.0);
    std::vector<cx<T>> scratch((size_t)grid * K::scratch_v());
    std::vector<unsigned long long> cnt(grid, 0);
    DpCsdParams<T> prm{};
    prm.traces = traces.data();
    prm.ev_stride = (long long)NCH * G::N;
    prm.chan_stride = G::N;
    prm.n_events = n_events;
    prm.mask = mask.data();
    prm.tw1 = dt.tw1.data();
    prm.tw2 = dt.tw2.data();
    prm.tw3 = dt.tw3.data();
    prm.twn = dt.twn.data();
    prm.groups = dt.groups.data();
    prm.chunk3 = dt.chunk3.data();
    prm.scratch = scratch.data();
    prm.scratch_per_cta = K::scratch_v();
    prm.partial = partial.data();
    prm.partial_per_cta = ppc;
    prm.count = cnt.data();
    prm.scale = scale;
    prm.subtract_first = subtract_first;
    for (int b = 0; b < grid; ++b) {
        std::vector<unsigned char> smem(K::SMEM_BYTES + 64);
        unsigned char* sp = smem.data();
        sp += (64 - (reinterpret_cast<uintptr_t>(sp) & 63)) & 63;
        run_cta(G::NT, b, grid, [&] { K::run(prm, sp); });
    }
    const int nbins = G::M + 1;
    sums.assign((size_t)K::NCOMP * nbins, 0.0);
    for (int comp = 0; comp < K::NCOMP; ++comp)
        for (int k = 0; k < nbins; ++k)
            for (int b = 0; b < grid; ++b) sums[(size_t)comp * nbins + k] += partial[(size_t)b * ppc + (size_t)loc[k] * K::NCP + comp];
    count = 0;
    for (int b = 0; b < grid; ++b) count += (double)cnt[b];
}

template <class T> static int main_t(std::ifstream& f, const char* outp, int N, int n_events, int n, bool f32) {
    double fs, rms;
    rd(f, &fs, 1);
    rd(f, &rms, 1);
    std::vector<unsigned char> mask(n_events);
    rd(f, mask.data(), mask.size());
    std::vector<double> traces((size_t)n_events * n * N);
    rd(f, traces.data(), traces.size());
    if (!f) { std::fprintf(stderr, "short input\n"); return 2; }
    const double scale = f32 ? std::exp2(-std::round(std::log2(rms))) : 1.0;
    std::vector<double> sums;
    double count = 0;
#define DP_EMU_NCH(R1_)                                                                                  \
    switch (n) {                                                                                         \
        case 2: run_all<T, R1_, 2>(fs, scale, f32 ? 1 : 0, traces, mask, n_events, sums, count); break;  \
        case 3: run_all<T, R1_, 3>(fs, scale, f32 ? 1 : 0, traces, mask, n_events, sums, count); break;  \
        case 4: run_all<T, R1_, 4>(fs, scale, f32 ? 1 : 0, traces, mask, n_events, sums, count); break;  \
        default: std::fprintf(stderr, "unsupported n_chan\n"); return 3;                                  \
    }
    switch (dpplan2::r1_of(N)) {
        case 2: DP_EMU_NCH(2) break;
        case 4: DP_EMU_NCH(4) break;
        case 8: DP_EMU_NCH(8) break;
        default: std::fprintf(stderr, "unsupported N\n"); return 3;
    }
    std::ofstream o(outp, std::ios::binary);
    o.write(reinterpret_cast<const char*>(sums.data()), sizeof(double) * sums.size());
    o.write(reinterpret_cast<const char*>(&count), sizeof(double));
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: emu_csd in out\n"); return 1; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        int32_t h[4];
        rd(f, h, 4);
        if (h[3]) return main_t<f2>(f, argv[2], h[0], h[1], h[2], true);
        return main_t<double>(f, argv[2], h[0], h[1], h[2], false);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 5;
    }
}
