// dependent-chain latencies (cycles per op, single warp) of the fp64 instructions the step kernels are built from
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void chain(double *out, long long *cyc, double a, double b, int sel)
{
    double x = a + threadIdx.x, y = b;
    int p = sel;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __dadd_rn(x, y);
        if (OP == 1) x = __dmul_rn(x, y);
        if (OP == 2) x = __fma_rn(x, y, y);
        if (OP == 3) { asm volatile("{\n\t.reg .pred q;\n\tsetp.lt.f64 q, %0, %1;\n\t@q mul.rn.f64 %0, %0, 0d3FF0000000000001;\n\t@!q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(x) : "d"(y)); }  // DSETP -> predicated op
        if (OP == 4) x = (double)__double2int_rz(x) + y;       // F2I + I2F + DADD
        if (OP == 5) x = ceil(x) + y;                          // FRND + DADD
        if (OP == 6) x = (double)(float)x + y;                 // F2F x2 + DADD
        if (OP == 7) { p = min(p, (int)(p * 3 + i) & 0xFF00FF80); }  // int chain
        if (OP == 8) { asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mul.rn.f64 %0, %1, 0d3FF0000000000000;\n\t}" : "+d"(x) : "d"(y), "r"(p)); x = __dadd_rn(x, y); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + p;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// throughput: W warps each with 8 independent chains
template <int OP> __global__ void thr(double *out, long long *cyc, double a, double b)
{
    double x[8];
    for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) x[j] = __dadd_rn(x[j], b);
            if (OP == 1) x[j] = __fma_rn(x[j], b, b);
            if (OP == 2) x[j] = (double)(float)x[j];
            if (OP == 3) x[j] = (double)__double2int_rz(x[j]);
        }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 8192); cudaMalloc(&cyc, 8);
    const char *names[] = {"DADD", "DMUL", "DFMA", "DSETP+pred op", "F2I+I2F+DADD", "FRND(ceil)+DADD", "F2F.f32+F2F.f64+DADD", "int min/imad/lop chain (3 ops)", "pred DMUL(mov)+DADD"};
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.0, 1.0000001, 1); chain<OP><<<1, 32>>>(out, cyc, 1.0, 1.0000001, 1); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-34s %.2f cycles per iteration\n", names[OP], (double)h / N);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    const char *tn[] = {"DADD", "DFMA", "F2F pair", "F2I+I2F"};
#define RUNT(OP, W) thr<OP><<<1, 32 * W>>>(out, cyc, 1.0, 1.0000001); thr<OP><<<1, 32 * W>>>(out, cyc, 1.0, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("throughput %-10s %2d warps/SM: %.2f cycles per warp-instruction per scheduler\n", tn[OP], W, (double)h / (N * 8.0 * W / 4.0));
    RUNT(0, 4) RUNT(0, 16) RUNT(1, 16) RUNT(2, 4) RUNT(2, 16) RUNT(3, 16)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
