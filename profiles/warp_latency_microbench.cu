// Single-warp dependent-chain latencies on sm_100a, the cost model behind the segmentation flood fill:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat profiles/warp_latency_microbench.cu && ./lat
// Measured on B200 (1.965 GHz): LDS 29 cycles, dependent ALU op ~6, MATCH.ANY with 9 distinct values 115, BALLOT 30,
// ATOMS.OR + LDS 47, STS/syncwarp/LDS 46, divergent if + syncwarp 38, STG + 2 ALU 39.  (Chains the compiler folds print 0.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, int* gbuf) {
  __shared__ int s[1024];
  __shared__ unsigned vis[256];
  const int lane = threadIdx.x;
  for (int i = lane; i < 1024; i += 32) s[i] = (i * 7 + 3) & 1023;
  for (int i = lane; i < 256; i += 32) vis[i] = 0;
  __syncwarp();
  const int N = 2000;
  long long t0, t1;
  int x = lane;
  // 0: dependent LDS chain
  t0 = clock64();
  for (int i = 0; i < N; ++i) x = s[x];
  t1 = clock64();
  if (lane == 0) out[0] = (t1 - t0) / N;
  // 1: dependent IADD chain (IMAD)
  int y = x;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { y = y * 3 + 1; y = y ^ (y >> 3); y = y + 7; y = y * 5; }
  t1 = clock64();
  if (lane == 0) out[1] = (t1 - t0) / (N * 4);
  // 2: match_any with 1 distinct + dependent
  unsigned m = y;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) m = __match_any_sync(0xffffffffu, (int)(m & 1));
  t1 = clock64();
  if (lane == 0) out[2] = (t1 - t0) / N;
  // 3: match_any with 9 distinct values
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) m = __match_any_sync(0xffffffffu, (int)((lane < 8 ? lane : -1) + (m & 0)));
  t1 = clock64();
  if (lane == 0) out[3] = (t1 - t0) / N;
  // 4: ballot dependent
  unsigned b = m;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) b = __ballot_sync(0xffffffffu, (b >> lane) & 1 || lane == i % 32);
  t1 = clock64();
  if (lane == 0) out[4] = (t1 - t0) / N;
  // 5: shfl dependent
  int z = b;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) z = __shfl_sync(0xffffffffu, z, (z + 1) & 31);
  t1 = clock64();
  if (lane == 0) out[5] = (t1 - t0) / N;
  // 6: ATOMS.OR (no return) followed by dependent LDS of same array
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { atomicOr(&vis[(z + lane) & 255], 1u << (i & 31)); z = vis[(z + 3 * lane + 1) & 255] & 255; }
  t1 = clock64();
  if (lane == 0) out[6] = (t1 - t0) / N;
  // 7: STS then syncwarp then LDS from another lane's slot
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { s[lane] = z; __syncwarp(); z = s[(lane + 1) & 31] + 1; __syncwarp(); }
  t1 = clock64();
  if (lane == 0) out[7] = (t1 - t0) / N;
  // 8: divergent branch + reconverge
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { if ((z + lane) & 1) z = z * 3 + 1; else z = z + 5; z &= 1023; __syncwarp(); }
  t1 = clock64();
  if (lane == 0) out[8] = (t1 - t0) / N;
  // 9: global store (fire and forget) per iteration + dependent ALU
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) { gbuf[(z & 1023) * 32 + lane] = z; z = z * 3 + 1; z &= 1023; }
  t1 = clock64();
  if (lane == 0) out[9] = (t1 - t0) / N;
  // 10: LDS.U8
  unsigned char* sb = reinterpret_cast<unsigned char*>(s);
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N; ++i) z = sb[(z * 4) & 4095] + (z & 0x300);
  t1 = clock64();
  if (lane == 0) out[10] = (t1 - t0) / N;
  if (z == 12345678 || m == 0x12345 || b == 0x54321) out[11] = x + y;
}
int main() {
  long long* d; int* g;
  cudaMalloc(&d, 16 * 8); cudaMalloc(&g, 1024 * 32 * 4);
  cudaMemset(d, 0, 128);
  k<<<1, 32>>>(d, g); cudaDeviceSynchronize();
  k<<<1, 32>>>(d, g);
  long long h[16]; cudaMemcpy(h, d, 128, cudaMemcpyDeviceToHost);
  const char* names[] = {"LDS chain", "ALU dep op", "MATCH 1 value", "MATCH 9 values", "BALLOT dep", "SHFL dep", "ATOMS.OR + LDS", "STS sync LDS sync", "divergent if + syncwarp", "STG + 2 ALU", "LDS.U8 chain"};
  for (int i = 0; i < 11; ++i) printf("%-24s %lld cycles\n", names[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
