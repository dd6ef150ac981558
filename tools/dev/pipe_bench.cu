// Micro-benchmark of issue rates on sm_100a: per-SMSP cycles per warp-instruction for a few opcodes and mixes.
// Development tool (not part of the library): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 16

template <int MODE>
__global__ void k(float* out, int* iout, long long* cyc, float seed) {
    const int one = (int)gridDim.y, mone = -(int)gridDim.z;
    float a[UNROLL], b = seed + threadIdx.x, c = seed * 0.5f;
    int ia[UNROLL], ib = (int)seed + threadIdx.x, ic = 3;
    float2 p[UNROLL / 2];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = seed + i; ia[i] = (int)seed + i * 7; }
#pragma unroll
    for (int i = 0; i < UNROLL / 2; ++i) p[i] = make_float2(seed + i, seed - i);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#define MN(i, d) a[i] = fminf(a[i], a[((i) + (d)) % UNROLL])
#define MX(i, d) a[i] = fmaxf(a[i], a[((i) + (d)) % UNROLL])
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {      // pass A
            if (MODE == 0) a[i] = fmaf(a[i], b, c);                                   // FFMA
            if (MODE == 1) MN(i, 5);                                                  // FMNMX
            if (MODE == 2) ia[i] = min(ia[i], ia[(i + 5) % UNROLL]);                  // VIMNMX
            if (MODE == 3) ia[i] = ia[i] * ib + ic;                                   // IMAD
            if (MODE == 4) ia[i] = __mulhi(ia[i], ib) + ic;                           // IMAD.HI
            if (MODE == 5) a[i] = a[i] + b;                                           // FADD
            if (MODE == 6) { if (i & 1) MN(i, 4); else a[i] = fmaf(a[i], b, c); }     // FMNMX + FFMA 1:1
            if (MODE == 7) { if (i & 1) MN(i, 4); else ia[i] = ia[i] * ib + ic; }     // FMNMX + IMAD 1:1
            if (MODE == 8 && i < UNROLL / 2) {                                        // FFMA2
                asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; mov.b64 z, {%3, %3};\n"
                             "fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x; }"
                             : "+f"(p[i].x), "+f"(p[i].y) : "f"(b), "f"(c));
            }
            if (MODE == 9 && i < UNROLL / 2) {                                        // FADD2
                asm volatile("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2};\n"
                             "add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
                             : "+f"(p[i].x), "+f"(p[i].y) : "f"(b));
            }
            if (MODE == 10) a[i] = fminf(fminf(a[i], a[(i + 5) % UNROLL]), a[(i + 9) % UNROLL]);     // FMNMX3
            if (MODE == 11) { if (i % 3 == 0) ia[i] = ia[i] * ib + ic; else MN(i, 6); }               // 2 FMNMX : 1 IMAD
            if (MODE == 12) { if (i % 4 == 0) ia[i] = ia[i] * ib + ic; else MN(i, 8); }               // 3 FMNMX : 1 IMAD
            if (MODE == 13) ia[i] = ia[i] + ia[(i + 5) % UNROLL] - ia[(i + 9) % UNROLL];               // IADD3
            if (MODE == 16) ia[i] = __vimin3_s32(ia[i], ia[(i + 5) % UNROLL], ia[(i + 9) % UNROLL]);   // VIMNMX3
            if (MODE == 17) ia[i] = (int)__vminu2((unsigned)ia[i], (unsigned)ia[(i + 5) % UNROLL]);    // packed u16x2 min
            if (MODE == 18) { __half2 h = __hmin2(*reinterpret_cast<__half2*>(&ia[i]), *reinterpret_cast<__half2*>(&ia[(i + 5) % UNROLL])); ia[i] = *reinterpret_cast<int*>(&h); }   // HMNMX2
            if (MODE == 19) ia[i] = (int)min((unsigned)ia[i], (unsigned)ia[(i + 5) % UNROLL]);         // unsigned min
            if (MODE == 14 && (i & 1) == 0) {   // compare-exchange, both FMNMX (2 instr per pair i, i+1)
                const float lo = fminf(a[i], a[i + 1]), hi = fmaxf(a[i], a[i + 1]); a[i] = lo; a[i + 1] = hi;
            }
            if (MODE == 15 && (i & 1) == 0) {   // compare-exchange: FMNMX + 2 IMAD with run-time +1 / -1 (3 instr per pair)
                const float lo = fminf(a[i], a[i + 1]);
                const int s2 = __float_as_int(a[i]) * one + __float_as_int(a[i + 1]);
                a[i + 1] = __int_as_float(__float_as_int(lo) * mone + s2); a[i] = lo;
            }
        }
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {      // pass B (max instead of min so that nothing can be merged)
            if (MODE == 0) a[i] = fmaf(a[i], c, b);
            if (MODE == 1) MX(i, 3);
            if (MODE == 2) ia[i] = max(ia[i], ia[(i + 3) % UNROLL]);
            if (MODE == 16) ia[i] = __vimax3_s32(ia[i], ia[(i + 3) % UNROLL], ia[(i + 7) % UNROLL]);
            if (MODE == 17) ia[i] = (int)__vmaxu2((unsigned)ia[i], (unsigned)ia[(i + 3) % UNROLL]);
            if (MODE == 18) { __half2 h = __hmax2(*reinterpret_cast<__half2*>(&ia[i]), *reinterpret_cast<__half2*>(&ia[(i + 3) % UNROLL])); ia[i] = *reinterpret_cast<int*>(&h); }
            if (MODE == 19) ia[i] = (int)max((unsigned)ia[i], (unsigned)ia[(i + 3) % UNROLL]);
            if (MODE == 3) ia[i] = ia[i] * ic + ib;
            if (MODE == 4) ia[i] = __mulhi(ia[i], ic) + ib;
            if (MODE == 5) a[i] = a[i] + c;
            if (MODE == 6) { if (i & 1) MX(i, 2); else a[i] = fmaf(a[i], c, b); }
            if (MODE == 7) { if (i & 1) MX(i, 2); else ia[i] = ia[i] * ic + ib; }
            if (MODE == 8 && i < UNROLL / 2) {
                asm volatile("{ .reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2}; mov.b64 z, {%3, %3};\n"
                             "fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x; }"
                             : "+f"(p[i].x), "+f"(p[i].y) : "f"(c), "f"(b));
            }
            if (MODE == 9 && i < UNROLL / 2) {
                asm volatile("{ .reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %2};\n"
                             "add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x; }"
                             : "+f"(p[i].x), "+f"(p[i].y) : "f"(c));
            }
            if (MODE == 10) a[i] = fmaxf(fmaxf(a[i], a[(i + 3) % UNROLL]), a[(i + 7) % UNROLL]);
            if (MODE == 11) { if (i % 3 == 0) ia[i] = ia[i] * ic + ib; else MX(i, 3); }
            if (MODE == 12) { if (i % 4 == 0) ia[i] = ia[i] * ic + ib; else MX(i, 4); }
            if (MODE == 13) ia[i] = ia[i] + ia[(i + 3) % UNROLL] - ia[(i + 7) % UNROLL];
            if (MODE == 14 && (i & 1) == 1 && i + 1 < UNROLL) {
                const float lo = fminf(a[i], a[i + 1]), hi = fmaxf(a[i], a[i + 1]); a[i] = lo; a[i + 1] = hi;
            }
            if (MODE == 15 && (i & 1) == 1 && i + 1 < UNROLL) {
                const float lo = fminf(a[i], a[i + 1]);
                const int s2 = __float_as_int(a[i]) * one + __float_as_int(a[i + 1]);
                a[i + 1] = __int_as_float(__float_as_int(lo) * mone + s2); a[i] = lo;
            }
        }
    }
    long long t1 = clock64();
    float s = 0; int is = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { s += a[i]; is += ia[i]; }
#pragma unroll
    for (int i = 0; i < UNROLL / 2; ++i) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = is;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps_per_smsp, int n_instr_per_iter) {
    float* out; int* iout; long long* cyc; long long h;
    const int threads = warps_per_smsp * 4 * 32;
    cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&iout, 148 * threads * 4); cudaMalloc(&cyc, 8);
    k<MODE><<<148, threads>>>(out, iout, cyc, 1.0001f);
    k<MODE><<<148, threads>>>(out, iout, cyc, 1.0001f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / ((double)ITERS * 2 * n_instr_per_iter * warps_per_smsp);
    printf("%-28s warps/SMSP=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, per);
    cudaFree(out); cudaFree(iout); cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4}) {
        run<0>("FFMA", w, UNROLL);
        run<5>("FADD", w, UNROLL);
        run<1>("FMNMX", w, UNROLL);
        run<10>("FMNMX3 (min3)", w, UNROLL);
        run<2>("IMNMX (int min)", w, UNROLL);
        run<19>("UMNMX (unsigned min)", w, UNROLL);
        run<16>("VIMNMX3 (int min3)", w, UNROLL);
        run<17>("VMNMX u16x2", w, UNROLL);
        run<18>("HMNMX2 (half2 min)", w, UNROLL);
        run<3>("IMAD", w, UNROLL);
        run<4>("IMAD.HI", w, UNROLL);
        run<6>("FMNMX+FFMA 1:1", w, UNROLL);
        run<7>("FMNMX+IMAD 1:1", w, UNROLL);
        run<11>("FMNMX+IMAD 2:1", w, UNROLL);
        run<12>("FMNMX+IMAD 3:1", w, UNROLL);
        run<13>("IADD3", w, UNROLL);
        run<14>("CE = 2 FMNMX   (per CE)", w, (2 * UNROLL - 1) / 4);
        run<15>("CE = FMNMX + 2 IMAD (per CE)", w, (2 * UNROLL - 1) / 4);
        run<8>("FFMA2 (f32x2)", w, UNROLL / 2);
        run<9>("FADD2 (f32x2)", w, UNROLL / 2);
    }
    return 0;
}
