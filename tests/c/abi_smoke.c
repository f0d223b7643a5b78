/* C translation unit OUTSIDE the library that includes include/birefnet_b200.h, links libbirefnet_b200.so and walks the
 * reference's call sequence through the C ABI -- what a Rust `-sys` crate, cgo or JNI stub would bind:
 *
 *   BiRefNet::new(config, vb)      ->  brn_model_create + brn_model_load_safetensors + brn_model_finalize
 *   model.forward_logits(&x)       ->  brn_forward_logits   (src/birefnet.rs:412-461)
 *   model.forward(&x)              ->  brn_forward          (src/birefnet.rs:466-469)
 *
 * Usage:  abi_smoke probe                                     (no GPU needed: config / version / error behaviour)
 *         abi_smoke run <weights.safetensors> <in.f32> <B> <H> <W> <embed> <h0> <h1> <h2> <h3> <out.f32>
 * `run` writes the logits as raw fp32; tests/test_abi_c.py compares them bit for bit with the ctypes path.
 * Compiled by the test with `gcc -std=c99 -Wall -Wextra -Werror`: a header that stops being valid C, or a prototype that
 * no longer matches the exported symbol's use here, fails the build. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "birefnet_b200.h"

static int fail(const char* what, brn_status st) {
  fprintf(stderr, "abi_smoke: %s failed with status %d: %s\n", what, (int)st, brn_last_error());
  return 1;
}

static int probe(void) {
  brn_config cfg;
  brn_config_swin_l(&cfg);
  if (cfg.embed_dim != 192 || cfg.depths[2] != 18 || cfg.num_heads[3] != 48 || cfg.window_size != 12) {
    fprintf(stderr, "abi_smoke: brn_config_swin_l does not match SwinConfig::swin_l (src/swin.rs:69-80)\n");
    return 1;
  }
  if (strstr(brn_version(), "sm_100a") == NULL) return 1;
  /* null arguments are a status, never a crash */
  if (brn_model_create(NULL, 0, NULL) != BRN_ERR_INVALID) return 1;
  if (brn_forward_logits(NULL, NULL, 1, 32, 32, 0, NULL, 0, NULL) == BRN_OK) return 1;
  brn_model* m = NULL;
  brn_status st = brn_model_create(&cfg, 0, &m);
  if (st == BRN_OK) {
    /* a GPU is present: call-order errors (candle's Result) */
    float x[3 * 32 * 32] = {0}, y[32 * 32];
    if (brn_forward_logits(m, x, 1, 32, 32, 0, y, 0, NULL) != BRN_ERR_STATE) return 1;
    if (brn_model_finalize(m) != BRN_ERR_MISSING_TENSOR) return 1;
    brn_model_destroy(m);
    printf("probe ok (GPU present)\n");
  } else {
    /* no GPU: the library must fail loudly, there is no CPU fallback */
    if (st != BRN_ERR_CUDA || strstr(brn_last_error(), "no CPU fallback") == NULL) return fail("brn_model_create", st);
    printf("probe ok (no GPU: %s)\n", brn_last_error());
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "probe") == 0) return probe();
  if (argc != 13 || strcmp(argv[1], "run") != 0) {
    fprintf(stderr, "usage: %s probe | run <weights> <in.f32> <B> <H> <W> <embed> <h0> <h1> <h2> <h3> <out.f32>\n", argv[0]);
    return 2;
  }
  const int B = atoi(argv[4]), H = atoi(argv[5]), W = atoi(argv[6]);
  brn_config cfg;
  brn_config_swin_l(&cfg);
  cfg.embed_dim = atoi(argv[7]);
  for (int i = 0; i < 4; ++i) { cfg.num_heads[i] = atoi(argv[8 + i]); cfg.depths[i] = 2; }
  cfg.precision = BRN_PREC_FP16;
  cfg.deform_mode = BRN_DEFORM_DEFORMABLE;

  brn_model* m = NULL;
  brn_status st = brn_model_create(&cfg, 0, &m);
  if (st != BRN_OK) return fail("brn_model_create", st);
  int32_t n_loaded = 0;
  if ((st = brn_model_load_safetensors(m, argv[2], &n_loaded)) != BRN_OK) return fail("brn_model_load_safetensors", st);
  if (n_loaded != brn_model_num_tensors(m)) { fprintf(stderr, "loaded %d of %d tensors\n", n_loaded, brn_model_num_tensors(m)); return 1; }
  if ((st = brn_model_finalize(m)) != BRN_OK) return fail("brn_model_finalize", st);

  const size_t nin = (size_t)B * 3 * H * W, nout = (size_t)B * H * W;
  float* x = (float*)brn_host_alloc(nin * sizeof(float));       /* pinned buffers from the library */
  float* y = (float*)brn_host_alloc(nout * sizeof(float));
  float* p = (float*)malloc(nout * sizeof(float));               /* and a pageable one */
  if (!x || !y || !p) return 1;
  FILE* f = fopen(argv[3], "rb");
  if (!f || fread(x, sizeof(float), nin, f) != nin) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
  fclose(f);
  if ((st = brn_forward_logits(m, x, B, H, W, 0, y, 0, NULL)) != BRN_OK) return fail("brn_forward_logits", st);
  if ((st = brn_forward(m, x, B, H, W, 0, p, 0, NULL)) != BRN_OK) return fail("brn_forward", st);
  for (size_t i = 0; i < nout; ++i) {
    const float s = 1.0f / (1.0f + expf(-y[i]));
    if (!(fabsf(s - p[i]) < 1e-5f)) { fprintf(stderr, "brn_forward != sigmoid(brn_forward_logits) at %zu\n", i); return 1; }
  }
  if (brn_launch_count(m) <= 0) { fprintf(stderr, "no kernels were launched\n"); return 1; }
  f = fopen(argv[12], "wb");
  if (!f || fwrite(y, sizeof(float), nout, f) != nout) return 1;
  fclose(f);
  brn_host_free(x); brn_host_free(y); free(p);
  brn_model_destroy(m);
  printf("run ok: %d tensors, %lld launches\n", (int)n_loaded, (long long)brn_launch_count(NULL));
  return 0;
}
