// C++ host-side smoke of the adapters (slam-eds_b200/host/edsgpu_adapters.hpp): reads one synthetic
// window written by tests/test_cpp_adapters.py, runs EventFrame::create + Tracker::optimize through
// the C ABI and prints the result for the test to compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../slam-eds_b200/host/edsgpu_adapters.hpp"

template <typename T>
static std::vector<T> rd(std::ifstream& f, size_t n) {
    std::vector<T> v(n);
    f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * n);
    return v;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: adapter_demo <input.bin>\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary);
    auto hdr = rd<int32_t>(f, 6);  // H W N E B iters
    const int H = hdr[0], W = hdr[1], N = hdr[2], E = hdr[3], B = hdr[4], iters = hdr[5];
    auto intr = rd<double>(f, 4);
    auto grad = rd<double>(f, 2 * (size_t)N), nc = rd<double>(f, 2 * (size_t)N), idp = rd<double>(f, N), wts = rd<double>(f, N);
    auto x0 = rd<double>(f, 13);
    auto ex = rd<uint16_t>(f, E), ey = rd<uint16_t>(f, E);
    auto ep = rd<uint8_t>(f, E);
    auto ts = rd<int64_t>(f, E);
    try {
        edsgpu_host::Context ctx(0);
        eds::tracking::EventFrame ef(ctx, (uint16_t)H, (uint16_t)W);
        std::vector<eds::tracking::Event> events(E);
        for (int i = 0; i < E; ++i) events[i] = {ex[i], ey[i], ts[i], ep[i]};
        ef.create(1, events);
        auto kf = std::make_shared<eds::tracking::KeyFrame>(ctx, grad, nc, idp, wts, H, W, intr[0], intr[1], intr[2], intr[3], B);
        eds::tracking::Config cfg;
        cfg.options.num_threads = B;
        cfg.options.max_num_iterations = {iters};
        cfg.loss_params = {0.05};
        eds::tracking::Tracker tracker(ctx, kf, cfg);
        std::array<double, 6> v0{{x0[7], x0[8], x0[9], x0[10], x0[11], x0[12]}};
        tracker.reset(kf, {{x0[0], x0[1], x0[2]}}, {{x0[3], x0[4], x0[5], x0[6]}}, &v0);
        std::array<double, 3> t{};
        std::array<double, 4> q{};
        const bool ok = tracker.optimize(0, ef, t, q);
        std::printf("ok %d\nnorm %.17g\ntime %lld\n", ok ? 1 : 0, ef.norm, (long long)ef.time_us);
        std::printf("px %.17g %.17g %.17g\nqx %.17g %.17g %.17g %.17g\n", tracker.px[0], tracker.px[1], tracker.px[2], tracker.qx[0], tracker.qx[1],
                    tracker.qx[2], tracker.qx[3]);
        std::printf("T_kf_ef %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", t[0], t[1], t[2], q[0], q[1], q[2], q[3]);
        std::printf("tau %.17g\niterations %d\nres0 %.17g\n", tracker.getLossParams()[0], tracker.getInfo().num_iterations, kf->residuals[0]);
        // the reference throws on non-monotonic event time (EventFrame.cpp:325-329)
        std::swap(events.front().ts_us, events.back().ts_us);
        if (events.front().ts_us > events.back().ts_us) {
            try { ef.create(2, events); std::printf("throw 0\n"); } catch (const std::runtime_error&) { std::printf("throw 1\n"); }
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
