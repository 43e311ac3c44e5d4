// Micro-benchmark of the host sweep (lps_sweep_votes; host only, no GPU) on the sweep inputs of a real contig:
//   LPS_DUMP_SWEEP=/tmp/sweep_in.bin python bench.py --steps 1 --warmup 1 --contigs-per-gpu 1      (on the GPU box; writes the dump)
//   g++ -O3 -std=c++17 -I../../include main.cpp -L../../longphase-s_b200 -l:liblps_b200.so -Wl,-rpath,$PWD/../../longphase-s_b200 \
//       -o /tmp/sweep_bench && LPS_SWEEP=avx512 /tmp/sweep_bench /tmp/sweep_in.bin
// The dump holds the rows as the device wrote them (block-shifted, header[3] bytes per row), which is what lps_host_sweep takes.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "lps.h"
// internal (C++ linkage, exported by the shared library): the sweep on the device's own row layout
int lps_host_sweep(const lps_phase_params *p, int32_t n_nodes, int32_t window, const int32_t *node_pos, const uint8_t *node_type,
                   const uint8_t *votes, const int8_t *last_link, int32_t *node_ps, int8_t *node_hap_ref);
int main(int argc, char **argv) {
    FILE *f = fopen(argc > 1 ? argv[1] : "/tmp/sweep_in.bin", "rb");
    if (!f) { perror("dump"); return 1; }
    int32_t hdr[4];
    if (fread(hdr, 4, 4, f) != 4) return 1;
    const int N = hdr[0], W = hdr[1], RS = hdr[3];
    std::vector<int32_t> pos(N), ps(N);
    std::vector<uint8_t> type(N), rowbuf((size_t)N * RS + 64);
    uint8_t *rows = rowbuf.data() + ((16 - ((uintptr_t)rowbuf.data() & 15)) & 15);
    std::vector<int8_t> last(N), hap(N);
    if (fread(pos.data(), 4, N, f) != (size_t)N || fread(type.data(), 1, N, f) != (size_t)N ||
        fread(rows, 1, (size_t)N * RS, f) != (size_t)N * RS || fread(last.data(), 1, N, f) != (size_t)N) return 1;
    fclose(f);
    lps_phase_params p{}; p.distance = hdr[2];
    for (int rep = 0; rep < 7; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        const int simd = lps_host_sweep(&p, N, W, pos.data(), type.data(), rows, last.data(), ps.data(), hap.data());
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        unsigned long cs = 0; for (int k = 0; k < N; k++) cs = cs * 1000003u + (unsigned long)(ps[k] * 3 + hap[k]);
        printf("sweep path %d: %.3f ms (%.1f ns/node) checksum %lx\n", simd, ms, ms * 1e6 / N, cs);
    }
}
