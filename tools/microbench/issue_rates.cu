// Issue cost (cycles per warp instruction) of the softmax instruction mix on sm_100a, for ONE warp alone on its scheduler
// and for TWO warps sharing a scheduler (warps 0 and 4 of the block). nvcc -arch=sm_100a -o issue_rates issue_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(x) x x x x x x x x
#define REP64(x) REP8(REP8(x))

template <int kOp>
__global__ void k(long long* out, float seed, int active_mask) {
  const int warp = threadIdx.x >> 5;
  if (!((active_mask >> warp) & 1)) return;
  float r[8];
  unsigned long long q[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
#pragma unroll
  for (int i = 0; i < 4; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q[i]) : "f"(r[2 * i]), "f"(r[2 * i + 1]));
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 16; ++it) {
    if (kOp == 0) {  // MUFU.EX2, 8 independent chains
      REP8(asm volatile("ex2.approx.ftz.f32 %0, %0;\n ex2.approx.ftz.f32 %1, %1;\n ex2.approx.ftz.f32 %2, %2;\n ex2.approx.ftz.f32 %3, %3;\n"
                        "ex2.approx.ftz.f32 %4, %4;\n ex2.approx.ftz.f32 %5, %5;\n ex2.approx.ftz.f32 %6, %6;\n ex2.approx.ftz.f32 %7, %7;"
                        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));)
    } else if (kOp == 1) {  // 3-input max
      REP8(asm volatile("max.f32 %0, %0, %1, %2;\n max.f32 %1, %1, %2, %3;\n max.f32 %2, %2, %3, %4;\n max.f32 %3, %3, %4, %5;\n"
                        "max.f32 %4, %4, %5, %6;\n max.f32 %5, %5, %6, %7;\n max.f32 %6, %6, %7, %0;\n max.f32 %7, %7, %0, %1;"
                        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));)
    } else if (kOp == 2) {  // 2-input max
      REP8(asm volatile("max.f32 %0, %0, %1;\n max.f32 %1, %1, %2;\n max.f32 %2, %2, %3;\n max.f32 %3, %3, %4;\n"
                        "max.f32 %4, %4, %5;\n max.f32 %5, %5, %6;\n max.f32 %6, %6, %7;\n max.f32 %7, %7, %0;"
                        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));)
    } else if (kOp == 3) {  // FFMA2
      REP8(asm volatile("fma.rn.f32x2 %0, %0, %1, %2;\n fma.rn.f32x2 %1, %1, %2, %3;\n fma.rn.f32x2 %2, %2, %3, %0;\n fma.rn.f32x2 %3, %3, %0, %1;\n"
                        "fma.rn.f32x2 %0, %0, %1, %2;\n fma.rn.f32x2 %1, %1, %2, %3;\n fma.rn.f32x2 %2, %2, %3, %0;\n fma.rn.f32x2 %3, %3, %0, %1;"
                        : "+l"(q[0]), "+l"(q[1]), "+l"(q[2]), "+l"(q[3]));)
    } else if (kOp == 4) {  // FADD2
      REP8(asm volatile("add.rn.f32x2 %0, %0, %1;\n add.rn.f32x2 %1, %1, %2;\n add.rn.f32x2 %2, %2, %3;\n add.rn.f32x2 %3, %3, %0;\n"
                        "add.rn.f32x2 %0, %0, %1;\n add.rn.f32x2 %1, %1, %2;\n add.rn.f32x2 %2, %2, %3;\n add.rn.f32x2 %3, %3, %0;"
                        : "+l"(q[0]), "+l"(q[1]), "+l"(q[2]), "+l"(q[3]));)
    } else if (kOp == 5) {  // F2FP.BF16 pack
      REP8(asm volatile("{.reg .b32 t0,t1,t2,t3,t4,t5,t6,t7;\n cvt.rn.bf16x2.f32 t0, %0, %1;\n cvt.rn.bf16x2.f32 t1, %1, %2;\n cvt.rn.bf16x2.f32 t2, %2, %3;\n cvt.rn.bf16x2.f32 t3, %3, %4;\n"
                        "cvt.rn.bf16x2.f32 t4, %4, %5;\n cvt.rn.bf16x2.f32 t5, %5, %6;\n cvt.rn.bf16x2.f32 t6, %6, %7;\n cvt.rn.bf16x2.f32 t7, %7, %0;\n"
                        "mov.b32 %0, t0; mov.b32 %1, t1; mov.b32 %2, t2; mov.b32 %3, t3; mov.b32 %4, t4; mov.b32 %5, t5; mov.b32 %6, t6; mov.b32 %7, t7;}"
                        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));)
    } else if (kOp == 6) {  // scalar FFMA
      REP8(asm volatile("fma.rn.f32 %0, %0, %1, %2;\n fma.rn.f32 %1, %1, %2, %3;\n fma.rn.f32 %2, %2, %3, %4;\n fma.rn.f32 %3, %3, %4, %5;\n"
                        "fma.rn.f32 %4, %4, %5, %6;\n fma.rn.f32 %5, %5, %6, %7;\n fma.rn.f32 %6, %6, %7, %0;\n fma.rn.f32 %7, %7, %0, %1;"
                        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]));)
    } else if (kOp == 7) {  // the softmax mix per 4 pairs: 8 MUFU + 4 FFMA2 + 4 FADD2 + 4 F2FP (20 instructions), independent of each other
      REP8(asm volatile("{.reg .b32 t0,t1,t2,t3;\n"
                        "ex2.approx.ftz.f32 %0, %0;\n ex2.approx.ftz.f32 %1, %1;\n fma.rn.f32x2 %8, %8, %9, %10;\n add.rn.f32x2 %9, %9, %10;\n cvt.rn.bf16x2.f32 t0, %4, %5;\n"
                        "ex2.approx.ftz.f32 %2, %2;\n ex2.approx.ftz.f32 %3, %3;\n fma.rn.f32x2 %10, %10, %11, %8;\n add.rn.f32x2 %11, %11, %8;\n cvt.rn.bf16x2.f32 t1, %5, %6;\n"
                        "ex2.approx.ftz.f32 %0, %0;\n ex2.approx.ftz.f32 %1, %1;\n fma.rn.f32x2 %8, %8, %9, %10;\n add.rn.f32x2 %9, %9, %10;\n cvt.rn.bf16x2.f32 t2, %6, %7;\n"
                        "ex2.approx.ftz.f32 %2, %2;\n ex2.approx.ftz.f32 %3, %3;\n fma.rn.f32x2 %10, %10, %11, %8;\n add.rn.f32x2 %11, %11, %8;\n cvt.rn.bf16x2.f32 t3, %7, %4;\n"
                        "mov.b32 %4, t0; mov.b32 %5, t1; mov.b32 %6, t2; mov.b32 %7, t3;}"
                        : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+l"(q[0]), "+l"(q[1]), "+l"(q[2]),
                          "+l"(q[3]));)
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += r[i];
  float lo, hi;
#pragma unroll
  for (int i = 0; i < 4; ++i) { asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[i])); acc += lo + hi; }
  if (threadIdx.x % 32 == 0) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = (long long)acc; }
}

template <int kOp>
void run(const char* name, int per_iter) {
  long long* d; cudaMalloc(&d, 64 * 8); long long h[64];
  for (int mask : {0x1, 0x11, 0x1111}) {
    cudaMemset(d, 0, 64 * 8);
    k<kOp><<<1, 32 * 13>>>(d, 0.5f, mask);
    k<kOp><<<1, 32 * 13>>>(d, 0.5f, mask);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps-on-scheduler %d : %6.2f cycles per warp instruction (warp 0: %lld cycles for %d)\n", name,
           mask == 0x1 ? 1 : mask == 0x11 ? 2 : 4, (double)h[0] / (16.0 * per_iter), h[0], 16 * per_iter);
  }
  cudaFree(d);
}

int main() {
  run<0>("MUFU.EX2", 64);
  run<1>("FMNMX3 (3-input max)", 64);
  run<2>("FMNMX (2-input max)", 64);
  run<3>("FFMA2 (fma.f32x2)", 64);
  run<4>("FADD2 (add.f32x2)", 64);
  run<5>("F2FP.BF16 pack (+MOV)", 128);
  run<6>("FFMA", 64);
  run<7>("mix 8 MUFU+4 FFMA2+4 FADD2+4 F2FP", 8 * 24);
  return 0;
}
