/* A plain-C caller of libhalva_dpa.so: proves that the boundary of include/halva_dpa.h is language-neutral (no Python, no
 * torch, no C++ types): dlopen the library, check the struct sizes, size the workspaces, run dpa_forward (fused-gradient
 * form) + dpa_backward on a batch read from a file, print the losses and gradient checksums.
 *   gcc smoke.c -I<repo>/include -I/usr/local/cuda/include -L/usr/local/cuda/lib64 -lcudart -ldl -lm -o smoke
 *   ./smoke <libhalva_dpa.so> <batch.bin>
 * batch.bin: int32[8] {B, T, B_ref, T_ref, H, V, 0, 0}, float alpha, float d_loss, then h_pair (bf16), labels_pair (int64),
 * signs_pair (int64), h_ref_pol (bf16), h_ref_frz (bf16), labels_ref (int64), W (bf16), each as raw little-endian arrays. */
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "halva_dpa.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static void* upload(FILE* f, size_t bytes) {
  void* h = malloc(bytes);
  void* d = NULL;
  if (fread(h, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(3); }
  if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { fprintf(stderr, "upload failed\n"); exit(3); }
  free(h);
  return d;
}
static double bf16_abs_sum(const void* dev, size_t n) {      /* checksum of a bf16 device array */
  uint16_t* h = (uint16_t*)malloc(n * 2);
  double s = 0.0;
  cudaMemcpy(h, dev, n * 2, cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < n; ++i) { uint32_t w = (uint32_t)h[i] << 16; float v; memcpy(&v, &w, 4); s += fabs((double)v); }
  free(h);
  return s;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: smoke <libhalva_dpa.so> <batch.bin>\n"); return 1; }
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  int (*version)(void) = (int (*)(void))dlsym(lib, "dpa_version");
  int (*sizes)(int*, int*, int*) = (int (*)(int*, int*, int*))dlsym(lib, "dpa_struct_sizes");
  int (*query)(const dpa_shape_t*, dpa_workspace_t*) = (int (*)(const dpa_shape_t*, dpa_workspace_t*))dlsym(lib, "dpa_query_workspace");
  int (*forward)(const dpa_args_t*, void*) = (int (*)(const dpa_args_t*, void*))dlsym(lib, "dpa_forward");
  int (*backward)(const dpa_args_t*, void*) = (int (*)(const dpa_args_t*, void*))dlsym(lib, "dpa_backward");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(lib, "dpa_last_error");
  if (!version || !sizes || !query || !forward || !backward || !last_error) { fprintf(stderr, "missing symbol\n"); return 1; }
  int sa, sb, sc;
  sizes(&sa, &sb, &sc);
  if (version() != DPA_ABI_VERSION || sa != (int)sizeof(dpa_shape_t) || sb != (int)sizeof(dpa_workspace_t) || sc != (int)sizeof(dpa_args_t)) {
    fprintf(stderr, "ABI mismatch: version %d, sizes %d %d %d\n", version(), sa, sb, sc);
    return 1;
  }
  FILE* f = fopen(argv[2], "rb");
  if (!f) { perror("batch"); return 1; }
  int32_t hd[8];
  float alpha, d_loss;
  if (fread(hd, 4, 8, f) != 8 || fread(&alpha, 4, 1, f) != 1 || fread(&d_loss, 4, 1, f) != 1) return 3;
  const int B = hd[0], T = hd[1], Br = hd[2], Tr = hd[3], H = hd[4], V = hd[5];
  dpa_args_t a;
  memset(&a, 0, sizeof(a));
  a.shape.B = B; a.shape.T = T; a.shape.B_ref = Br; a.shape.T_ref = Tr; a.shape.H = H; a.shape.V = V; a.shape.V_local = V;
  a.shape.dtype = DPA_BF16; a.shape.index_dtype = DPA_I64;
  a.shape.flags = DPA_FLAG_ROUND_LOGITS | DPA_FLAG_NEED_DW | DPA_FLAG_FUSED_GRAD;
  a.alpha = alpha;
  const size_t n_hp = (size_t)2 * B * T * H, n_hr = (size_t)Br * Tr * H, n_w = (size_t)V * H;
  a.h_pair = upload(f, n_hp * 2);
  a.labels_pair = upload(f, (size_t)2 * B * T * 8);
  a.signs_pair = upload(f, (size_t)2 * B * T * 8);
  a.h_ref_pol = upload(f, n_hr * 2);
  a.h_ref_frz = upload(f, n_hr * 2);
  a.labels_ref = upload(f, (size_t)Br * Tr * 8);
  a.W = upload(f, n_w * 2);
  a.W_ref = a.W;
  fclose(f);
  dpa_workspace_t ws;
  if (query(&a.shape, &ws)) { fprintf(stderr, "query: %s\n", last_error()); return 1; }
  CK(cudaMalloc(&a.state, ws.state_bytes)); a.state_bytes = ws.state_bytes;
  CK(cudaMalloc(&a.scratch, ws.fused_bytes)); a.scratch_bytes = ws.fused_bytes;
  CK(cudaMalloc((void**)&a.losses, 16));
  CK(cudaMalloc((void**)&a.status, 32));
  CK(cudaMalloc(&a.dH_pair, n_hp * 2));
  CK(cudaMalloc(&a.dH_ref, n_hr * 2));
  CK(cudaMalloc(&a.dW, n_w * 2));
  float* d_go = NULL;
  CK(cudaMalloc((void**)&d_go, 4));
  CK(cudaMemcpy(d_go, &d_loss, 4, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  if (forward(&a, st)) { fprintf(stderr, "dpa_forward: %s\n", last_error()); return 1; }
  a.d_loss = d_go;
  if (backward(&a, st)) { fprintf(stderr, "dpa_backward: %s\n", last_error()); return 1; }
  CK(cudaStreamSynchronize(st));
  float losses[4];
  int32_t status[8];
  CK(cudaMemcpy(losses, a.losses, 16, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(status, a.status, 32, cudaMemcpyDeviceToHost));
  printf("status %d K %d rows %d %d\n", status[0], status[1], status[2], status[3]);
  printf("loss %.9g align %.9g div %.9g\n", losses[0], losses[1], losses[2]);
  printf("dH_pair %.17g dH_ref %.17g dW %.17g\n", bf16_abs_sum(a.dH_pair, n_hp), bf16_abs_sum(a.dH_ref, n_hr), bf16_abs_sum(a.dW, n_w));
  return 0;
}
