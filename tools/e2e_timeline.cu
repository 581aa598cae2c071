// e2e_timeline.cu -- where the time of a host-buffer compress goes (tuning aid, not product).
// Rebuilds gpuar_b200_compress_host's pipeline on top of the public C ABI with timing events on
// every chunk, for a sweep of chunk counts:  H2D(chunk) -> encode -> [host learns the size] -> D2H.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude tools/e2e_timeline.cu \
//        -Lgpuar_b200 -lgpuar_b200 -Xlinker -rpath=/root/repo/gpuar_b200 -o gpuar_b200/csrc/build/e2e_timeline
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "gpuar_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

int main(int argc, char **argv)
{
    const size_t n = (argc > 1 ? atol(argv[1]) : 64) << 20;
    gpuar_b200_init();
    uint8_t *h_in, *h_out;
    CK(cudaMallocHost(&h_in, n));
    CK(cudaMallocHost(&h_out, gpuar_b200_payload_bound(n) + 64));
    uint64_t x = 0x64;
    for (size_t i = 0; i < n / 8; ++i) { x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; ((uint64_t *)h_in)[i] = z ^ (z >> 31); }
    const int kMax = 32;
    cudaStream_t st[kMax];
    cudaEvent_t e_h2d[kMax], e_enc[kMax], e_d2h[kMax], e0;
    uint8_t *d_in[kMax], *d_pay[kMax];
    void *d_scr[kMax];
    uint64_t *h_total;
    CK(cudaMallocHost(&h_total, kMax * 8));
    CK(cudaEventCreate(&e0));
    const size_t max_chunk = n / 2 + 8192;
    for (int l = 0; l < kMax; ++l) {
        CK(cudaStreamCreateWithFlags(&st[l], cudaStreamNonBlocking));
        CK(cudaEventCreate(&e_h2d[l])); CK(cudaEventCreate(&e_enc[l])); CK(cudaEventCreate(&e_d2h[l]));
    }
    cudaStream_t up, down;                     // ordered variant: every H2D on one stream, every D2H on another
    CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking));
    for (int ordered = 0; ordered < 2; ++ordered)
    for (int chunks : {2, 4, 8, 12, 16, 24, 32}) {
        size_t chunk = ((n / chunks + 8191) / 8192) * 8192;
        if (chunk > max_chunk) continue;
        for (int l = 0; l < chunks; ++l) {
            CK(cudaMalloc(&d_in[l], chunk + 64)); CK(cudaMalloc(&d_pay[l], gpuar_b200_payload_bound(chunk) + 64));
            CK(cudaMalloc(&d_scr[l], gpuar_b200_encode_scratch_bytes(chunk)));
        }
        double best = 1e9;
        std::vector<float> t_h2d(chunks), t_enc(chunks), t_d2h(chunks);
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaDeviceSynchronize());
            auto w0 = std::chrono::steady_clock::now();
            CK(cudaEventRecord(e0, st[0]));
            size_t pos = 20;
            for (int k = 0; k < chunks; ++k) {
                const size_t off = (size_t)k * chunk, m = n - off < chunk ? n - off : chunk;
                if (ordered) {
                    CK(cudaMemcpyAsync(d_in[k], h_in + off, m, cudaMemcpyHostToDevice, up));
                    CK(cudaEventRecord(e_h2d[k], up));
                    CK(cudaStreamWaitEvent(st[k], e_h2d[k], 0));
                } else {
                    CK(cudaMemcpyAsync(d_in[k], h_in + off, m, cudaMemcpyHostToDevice, st[k]));
                    CK(cudaEventRecord(e_h2d[k], st[k]));
                }
                gpuar_b200_encode(d_in[k], m, d_pay[k], gpuar_b200_payload_bound(chunk) + 64, &h_total[k], nullptr, d_scr[k],
                                  gpuar_b200_encode_scratch_bytes(chunk), st[k]);
                CK(cudaEventRecord(e_enc[k], st[k]));
            }
            for (int k = 0; k < chunks; ++k) {
                CK(cudaEventSynchronize(e_enc[k]));
                CK(cudaMemcpyAsync(h_out + pos, d_pay[k], h_total[k], cudaMemcpyDeviceToHost, ordered ? down : st[k]));
                CK(cudaEventRecord(e_d2h[k], ordered ? down : st[k]));
                pos += h_total[k];
            }
            for (int k = 0; k < chunks; ++k) CK(cudaStreamSynchronize(st[k]));
            CK(cudaStreamSynchronize(down));
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
            if (ms < best) {
                best = ms;
                for (int k = 0; k < chunks; ++k) {
                    CK(cudaEventElapsedTime(&t_h2d[k], e0, e_h2d[k])); CK(cudaEventElapsedTime(&t_enc[k], e0, e_enc[k]));
                    CK(cudaEventElapsedTime(&t_d2h[k], e0, e_d2h[k]));
                }
            }
        }
        printf("%s chunks %2d (%5.1f MiB): %.3f ms = %.1f GB/s | per chunk h2d/enc/d2h done at:", ordered ? "ordered  " : "per-lane ", chunks, chunk / 1048576.0, best, n / best / 1e6);
        for (int k = 0; k < chunks; ++k) if (k < 3 || k >= chunks - 2) printf(" [%d] %.2f/%.2f/%.2f", k, t_h2d[k], t_enc[k], t_d2h[k]);
        printf("\n");
        for (int l = 0; l < chunks; ++l) { cudaFree(d_in[l]); cudaFree(d_pay[l]); cudaFree(d_scr[l]); }
    }
    // reference points: one plain H2D and one plain D2H of the whole buffer
    uint8_t *d_all;
    CK(cudaMalloc(&d_all, n));
    for (int dir = 0; dir < 2; ++dir) {
        double best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaDeviceSynchronize());
            auto w0 = std::chrono::steady_clock::now();
            if (dir == 0) CK(cudaMemcpyAsync(d_all, h_in, n, cudaMemcpyHostToDevice, st[0]));
            else CK(cudaMemcpyAsync(h_out, d_all, n, cudaMemcpyDeviceToHost, st[0]));
            CK(cudaStreamSynchronize(st[0]));
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
            if (ms < best) best = ms;
        }
        printf("%s of %zu MiB: %.3f ms = %.1f GB/s\n", dir ? "D2H" : "H2D", n >> 20, best, n / best / 1e6);
    }
    return 0;
}
