// Host-thread emulation of ONE CTA of the mixed-radix OF kernel (dp_ofg_kernel.cuh; trace lengths that are not 2^k).
// Test infrastructure only.  Same input file format as emu_of.cpp.
// usage: emu_ofg <in.bin> <out.bin> <f32|f64>
#define DP_HOST_EMU 1
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../detprocess_b200/csrc/dp_ofg_plan.hpp"

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

template <class T> static int main_t(const char* in, const char* outp) {
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
    const int n_out = 1 + DP_SLOT_NOUT * n_fits;
    std::vector<double> out((size_t)n_events * n_out, -1.0);
    auto dt = dpgen::build_tables<T>(N, fs, chans, fcut, scale);
    DpGenChanDev<T> d{};
    d.wj_k = dt.chans[0].wj_k.data();
    d.wj_m = dt.chans[0].wj_m.data();
    d.wj_low = dt.chans[0].wj_low.data();
    d.adc_gain = 1.0;
    d.n_templ = n_templ;
    d.n_slots = n_fits;
    d.out_base = 0;
    for (int i = 0; i < n_templ; ++i) {
        auto& h = dt.chans[0].templ[i];
        d.templ[i].phi_k = h.phi_k.data();
        d.templ[i].phi_m = h.phi_m.data();
        d.templ[i].s_low = h.s_low.data();
        d.templ[i].norm = h.norm;
        d.templ[i].tsum = h.tsum;
        d.templ[i].pretrigger = h.pretrigger;
    }
    for (int i = 0; i < n_fits; ++i) d.slots[i] = DpSlot{ch.fits[i].templ, ch.fits[i].lo, ch.fits[i].hi, ch.fits[i].outside, dt.nlow};
    DpGenParams<T> prm{};
    prm.traces = traces.data();
    prm.in_dtype = 0;
    prm.event_stride = N;
    prm.chan_stride = N;
    prm.n_rows = n_events;
    prm.n_chan = 1;
    prm.chans = &d;
    prm.M = N / 2;
    prm.n_pass = (int)dt.radix.size();
    for (int j = 0; j < prm.n_pass; ++j) prm.radix[j] = dt.radix[j];
    prm.tw = dt.tw.data();
    prm.wn = dt.wn.data();
    prm.pos_k = dt.pos_k.data();
    prm.pos_m = dt.pos_m.data();
    prm.n_pairs = dt.n_pairs;
    prm.out = out.data();
    prm.n_out = n_out;
    prm.nlow = dt.nlow;
    prm.scale = dt.scale;
    prm.subtract_first = subtract_first;
    const int grid = 2;
    for (int b = 0; b < grid; ++b) {
        std::vector<unsigned char> smem(DpGenKernel<T>::smem_bytes(prm.M) + 64);
        unsigned char* sp = smem.data();
        sp += (64 - (reinterpret_cast<uintptr_t>(sp) & 63)) & 63;
        run_cta(DPG_NT, b, grid, [&] { DpGenKernel<T>::run(prm, sp); });
    }
    std::ofstream o(outp, std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: emu_ofg in out f32|f64\n"); return 1; }
    try {
        if (std::string(argv[3]) == "f32") return main_t<float>(argv[1], argv[2]);
        return main_t<double>(argv[1], argv[2]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 5;
    }
}
