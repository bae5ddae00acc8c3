// Host-thread emulation of ONE CTA of the fused OF kernel (see dp_platform.cuh).
// Test infrastructure only: checks index maths and barrier placement without a GPU.
// usage: emu_of <in.bin> <out.bin> <f32|f64>
#define DP_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../detprocess_b200/csrc/dp_plan.hpp"

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

template <class T, int R1, int P>
static void run_all(const dpplan::Geometry& g, dpplan::DeviceTables<T>& dt, const std::vector<dpplan::Channel>& chans,
                    const std::vector<double>& traces, int n_events, int subtract_first, std::vector<double>& out, int n_out) {
    using K = DpOfKernel<T, R1, P, 0>;
    std::vector<DpChanDev<T>> cd(chans.size());
    int base = 0;
    for (size_t c = 0; c < chans.size(); ++c) {
        auto& d = cd[c];
        d.wj = dt.chans[c].wj.data();
        d.wj_low = dt.chans[c].wj_low.data();
        d.wj_self = dt.chans[c].wj_self.data();
        d.n_templ = (int)chans[c].templ.size();
        d.n_slots = (int)chans[c].fits.size();
        d.out_base = base;
        base += 1 + DP_SLOT_NOUT * d.n_slots;
        for (int i = 0; i < d.n_templ; ++i) {
            auto& t = d.templ[i];
            auto& h = dt.chans[c].templ[i];
            t.phi = h.phi.data();
            t.phi_self = h.phi_self.data();
            t.s_low = h.s_low.data();
            t.norm = h.norm;
            t.tsum = h.tsum;
            t.pretrigger = h.pretrigger;
        }
        for (int i = 0; i < d.n_slots; ++i) d.slots[i] = DpSlot{chans[c].fits[i].templ, chans[c].fits[i].lo, chans[c].fits[i].hi, chans[c].fits[i].outside, dt.nlow};
    }
    const int grid = 2;
    std::vector<cx<T>> scratch((size_t)grid * 96 * g.NT);
    DpOfParams<T> prm{};
    prm.traces = traces.data();
    prm.event_stride = (long long)g.N * (long long)chans.size();
    prm.chan_stride = g.N;
    prm.n_rows = n_events * (int)chans.size();
    prm.n_chan = (int)chans.size();
    prm.chans = cd.data();
    prm.tw1 = dt.tw1.data();
    prm.tw2 = dt.tw2.data();
    prm.twn = dt.twn.data();
    prm.twp = dt.twp.data();
    prm.scratch = scratch.data();
    prm.scratch_per_cta = 96 * g.NT;
    prm.out = out.data();
    prm.n_out = n_out;
    prm.nlow = dt.nlow;
    prm.scale = dt.scale;
    prm.subtract_first = subtract_first;
    prm.in_dtype = 0;
    for (int b = 0; b < grid; ++b) {
        std::vector<unsigned char> smem(K::SMEM_BYTES + 64);
        unsigned char* sp = smem.data();
        sp += (64 - (reinterpret_cast<uintptr_t>(sp) & 63)) & 63;
        run_cta(g.NT, b, grid, [&] { K::run(prm, sp); });
    }
}

template <class T> static int main_t(const char* in, const char* outp, bool force_p2) {
    std::ifstream f(in, std::ios::binary);
    int32_t hdr[6];
    rd(f, hdr, 6);
    const int N = hdr[0], n_events = hdr[1], n_templ = hdr[2], n_fits = hdr[3], ac = hdr[4], subtract_first = hdr[5];
    double fs, fcut, scale;
    rd(f, &fs, 1);
    rd(f, &fcut, 1);
    rd(f, &scale, 1);
    std::vector<dpplan::Channel> chans(1);
    auto& ch = chans[0];
    ch.J.resize(N);
    rd(f, ch.J.data(), N);
    if (ac) ch.J[0] = std::numeric_limits<double>::infinity();
    for (int i = 0; i < n_templ; ++i) {
        dpplan::Template tp;
        int32_t h2[2];
        rd(f, h2, 2);
        tp.pretrigger = h2[0];
        tp.integralnorm = h2[1] != 0;
        tp.trace.resize(N);
        rd(f, tp.trace.data(), N);
        dpplan::finalize_template(tp, ch.J, fs);
        ch.templ.push_back(std::move(tp));
    }
    for (int i = 0; i < n_fits; ++i) {
        int32_t h4[4];
        rd(f, h4, 4);
        ch.fits.push_back(dpplan::Fit{h4[0], h4[1], h4[2], h4[3]});
    }
    std::vector<double> traces((size_t)n_events * N);
    rd(f, traces.data(), traces.size());
    if (!f) { std::fprintf(stderr, "short input\n"); return 2; }

    const bool f64 = sizeof(T) == 8;
    dpplan::Geometry g = dpplan::pick_geometry(N, f64);
    if (force_p2 && g.P == 1 && g.R1 >= 4) {
        g.P = 2;
        g.MS /= 2;
        g.R1 /= 2;
        g.NT /= 2;
    }
    const int n_out = 1 + DP_SLOT_NOUT * n_fits;
    std::vector<double> out((size_t)n_events * n_out, -1.0);
    auto dt = dpplan::build_tables<T>(g, fs, chans, fcut, scale);
    if (g.P == 1) {
        switch (g.R1) {
#define CASE(r) case r: run_all<T, r, 1>(g, dt, chans, traces, n_events, subtract_first, out, n_out); break;
            CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
            default: std::fprintf(stderr, "bad R1\n"); return 3;
        }
    } else {
        // the emulator also runs small split geometries the library does not build (P = 2, any R1)
        switch (g.R1) {
#define CASE(r) case r: run_all<T, r, 2>(g, dt, chans, traces, n_events, subtract_first, out, n_out); break;
            CASE(2) CASE(4) CASE(8) CASE(16) CASE(32)
#undef CASE
            default: std::fprintf(stderr, "bad R1\n"); return 3;
        }
    }
    std::ofstream o(outp, std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: emu_of in out f32|f64\n"); return 1; }
    try {
        const bool p2 = argc > 4 && std::string(argv[4]) == "p2";
        if (std::string(argv[3]) == "f32") return main_t<float>(argv[1], argv[2], p2);
        return main_t<double>(argv[1], argv[2], p2);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 5;
    }
}
