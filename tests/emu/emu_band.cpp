// Host-thread emulation of the band-amplitude kernel (test infrastructure only).
// usage: emu_band <in.bin> <out.bin>
// in:  int32 N, n_events, n_bands, in_dtype, stride | double fs, gain, offset | int32 lo[n_bands], hi[n_bands] | samples
#define DP_HOST_EMU 1
#include <cstdio>
#include <fstream>

#include "../../detprocess_b200/csrc/dp_band_kernel.cuh"

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
    int32_t hdr[5];
    rd(f, hdr, 5);
    const int N = hdr[0], n_events = hdr[1], n_bands = hdr[2], in_dtype = hdr[3], stride = hdr[4];
    double par[3];
    rd(f, par, 3);
    std::vector<int> lo(n_bands), hi(n_bands);
    rd(f, lo.data(), n_bands);
    rd(f, hi.data(), n_bands);
    const size_t esz = in_dtype == 0 ? 8 : (in_dtype == 1 ? 4 : 2);
    std::vector<unsigned char> samples((size_t)n_events * stride * esz);
    rd(f, samples.data(), samples.size());
    std::vector<double2> roots(N);
    for (int j = 0; j < N; ++j) {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)N;
        roots[j] = double2{(double)cosl(ang), (double)sinl(ang)};
    }
    std::vector<double> out((size_t)n_events * n_bands, -1.0);
    DpBandParams prm{};
    prm.base = samples.data();
    prm.in_dtype = in_dtype;
    prm.n_events = n_events;
    prm.event_stride = stride;
    prm.N = N;
    prm.gain = par[1];
    prm.offset = par[2];
    prm.bin_lo = lo.data();
    prm.bin_hi = hi.data();
    prm.n_bands = n_bands;
    prm.roots = roots.data();
    prm.norm = (double)N / (par[0] * par[0] * par[0]);
    prm.out = out.data();
    const int grid = 2;
    for (int b = 0; b < grid; ++b) {
        std::vector<double> s_re(DP_BAND_KB * DP_BAND_NT / 32), s_im(DP_BAND_KB * DP_BAND_NT / 32);
        double s_sum = 0.0;
        run_cta(DP_BAND_NT, b, grid, [&] { dp_band_rows(prm, s_re.data(), s_im.data(), &s_sum); });
    }
    std::ofstream o(argv[2], std::ios::binary);
    o.write(reinterpret_cast<const char*>(out.data()), sizeof(double) * out.size());
    return 0;
}
