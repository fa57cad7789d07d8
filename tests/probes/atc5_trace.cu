// Development harness: timeline of one attention_tc05 CTA in the middle of a full grid (B = 1024, L = 784, E = 64).
#define ATC5_TRACE 1
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../../controlnet-pytorch_b200/csrc/attention_tc05.cu"
namespace cnb {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void count_launch(int) {}
}
int main(int argc, char** argv) {
  const int B = 1024, L = 784, E = 64, heads = 4;
  const size_t n = (size_t)B * L * 3 * E;
  std::vector<__half> h(n);
  unsigned s = 12345;
  for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = __float2half(((s >> 8) & 0xFFFF) / 32768.0f - 1.0f); }
  __half *qkv, *out;
  cudaMalloc(&qkv, n * 2); cudaMalloc(&out, (size_t)B * L * E * 2);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  for (int it = 0; it < 2; ++it) cnb::attention_tc05(qkv, out, B, L, E, heads, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("err=%d flag=%d\n", (int)e, cnb::attention_tc05_error_flag());
  static long long tr[8][512];
  cudaMemcpyFromSymbol(tr, cnb::atc5::g_trace, sizeof(tr));
  long long t0 = 0;
  for (int r = 0; r < 6; ++r) for (int i = 0; i < 512; i += 2) if (tr[r][i + 1] && (!t0 || tr[r][i + 1] < t0)) t0 = tr[r][i + 1];
  const char* names[6] = {"S-issuer", "PV-issuer", "softmax hf0", "softmax hf1", "producer", "transposer"};
  for (int r = 0; r < 6; ++r) {
    printf("== %s\n", names[r]);
    for (int i = 0; i < 512 && tr[r][i + 1]; i += 2) printf("%lld:%lld ", tr[r][i], tr[r][i + 1] - t0);
    printf("\n");
  }
  return 0;
}
