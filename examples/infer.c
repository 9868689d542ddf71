/* Minimal C host for the ishara_b200 C ABI (include/ishara_b200.h): builds the BASELINE model, fills it with
 * deterministic weights, runs one batch through ishara_model_infer_host and prints the decoded strings.
 *   gcc -O2 -Iinclude examples/infer.c -Lishara_b200/lib -lishara_b200 -Wl,-rpath,$PWD/ishara_b200/lib -lm -o /tmp/infer
 * Without a Blackwell GPU every compute call fails with ISHARA_ERR_CUDA and a message: there is no CPU fallback. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ishara_b200.h"

#define CHECK(call)                                                               \
  do {                                                                            \
    ishara_status_t st_ = (call);                                                 \
    if (st_ != ISHARA_OK) {                                                       \
      fprintf(stderr, "%s -> status %d: %s\n", #call, (int)st_, ishara_last_error()); \
      return st_ == ISHARA_ERR_CUDA ? 3 : 1;                                      \
    }                                                                             \
  } while (0)

static float lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return (float)(*s >> 8) / 16777216.0f - 0.5f; }

int main(void) {
  ishara_config_t cfg = {256, 2, 2, 3, {11, 5, 3}, 3, 8, 2, 15, 384, 276, 60};  /* get_model(...) c7:1-11 */
  ishara_model_t* m = NULL;
  printf("%s, %d CUDA device(s)\n", ishara_version(), ishara_device_count());
  CHECK(ishara_model_create(&cfg, 0, &m));
  unsigned seed = 1;
  int np = ishara_model_num_params(m);
  long long total = 0;
  for (int i = 0; i < np; ++i) {
    const char* name;
    int64_t numel, shape[4];
    int32_t ndim;
    CHECK(ishara_model_param_info(m, i, &name, &numel, &ndim, shape));
    float* w = (float*)malloc(sizeof(float) * (size_t)numel);
    int is_var = 0, is_gamma = 0;
    for (const char* p = name; *p; ++p) {
      if (p[0] == 'v' && p[1] == 'a' && p[2] == 'r') is_var = 1;
      if (p[0] == 'g' && p[1] == 'a' && p[2] == 'm') is_gamma = 1;
    }
    for (int64_t j = 0; j < numel; ++j) w[j] = (is_var || is_gamma) ? 1.0f : 0.1f * lcg(&seed);
    CHECK(ishara_model_set_param(m, name, w, numel));
    free(w);
    total += numel;
  }
  printf("%d tensors, %lld parameters\n", np, total);  /* 7,591,096 for the BASELINE configuration */
  const int B = 2, T = cfg.frames, F = cfg.features;
  float* x = (float*)malloc(sizeof(float) * B * T * F);
  for (int i = 0; i < B * T * F; ++i) x[i] = 2.0f * lcg(&seed);
  int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * B * T);
  int32_t lens[2];
  CHECK(ishara_model_finalize(m));
  CHECK(ishara_model_infer_host(m, x, B, NULL, 0, NULL, ids, lens, NULL));
  static const char chars[] = " !#$%&'()*+,-./0123456789:;=?@[_abcdefghijklmnopqrstuvwxyz~";
  char* text = (char*)malloc((size_t)B * T + 1);
  int64_t offs[3];
  CHECK(ishara_ids_to_text(ids, lens, B, T, chars, 59, text, offs));
  for (int b = 0; b < B; ++b) printf("sequence %d: %d tokens: \"%.*s\"\n", b, lens[b], (int)(offs[b + 1] - offs[b]), text + offs[b]);
  CHECK(ishara_model_destroy(m));
  free(x); free(ids); free(text);
  return 0;
}
