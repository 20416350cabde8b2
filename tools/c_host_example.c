/* A host that is not Python: scores molecules with the reference's MixedInputModel (20250113.py:69-119) through the
 * whole-model C entry points of include/bbbp_b200.h.  Plain C99 + the CUDA runtime API for memory; no torch anywhere.
 *
 *   gcc -O2 -std=c99 tools/c_host_example.c -Iinclude -I/usr/local/cuda/include \
 *       -Lbbbp-multi-modal-deep-ensemble-framework_b200 -lbbbp_b200 -L/usr/local/cuda/lib64 -lcudart -o c_host_example
 *   c_host_example params.bin inputs.bin scores.bin <fingerprint_size> <groups> <seq> <precision 1|2|3> <image_is_u8 0|1>
 *
 * params.bin  the fp32 tensors of model.state_dict() back to back, in state_dict order (bbbp_model_param_name /
 *             bbbp_model_param_numel enumerate them; fc.2.num_batches_tracked is stored as one float and ignored)
 * inputs.bin  groups*seq fingerprint rows (fp32), then groups*seq images (3*128*128 fp32 values, or raw bytes)
 * scores.bin  groups*seq fp32 scores (written)
 * tests/test_model_gpu.py::test_c_program_scores_equal_the_python_host builds and runs this against the Python host. */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bbbp_b200.h"

#define CUDA_OK(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                          \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)
#define BBBP_OK_OR_DIE(call)                                                               \
  do {                                                                                     \
    int rc_ = (call);                                                                      \
    if (rc_ != BBBP_OK) {                                                                  \
      fprintf(stderr, "%s: status %d: %s\n", #call, rc_, bbbp_last_error());               \
      return 3;                                                                            \
    }                                                                                      \
  } while (0)

static void* read_file(const char* path, size_t* bytes) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  *bytes = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  void* buf = malloc(*bytes ? *bytes : 1);
  if (buf && fread(buf, 1, *bytes, f) != *bytes) {
    free(buf);
    buf = NULL;
  }
  fclose(f);
  return buf;
}

int main(int argc, char** argv) {
  if (argc != 9) {
    fprintf(stderr, "usage: %s params.bin inputs.bin scores.bin F groups seq precision image_is_u8\n", argv[0]);
    return 1;
  }
  bbbp_model_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.abi_version = BBBP_ABI_VERSION;
  desc.variant = BBBP_MODEL_TCNN_20250113;
  desc.fingerprint_size = atoi(argv[4]);
  desc.groups = atoi(argv[5]);
  desc.seq = atoi(argv[6]);
  desc.precision = atoi(argv[7]);
  desc.image_is_u8 = atoi(argv[8]);
  BBBP_OK_OR_DIE(bbbp_device_check());
  const int n_params = bbbp_model_param_count(&desc);
  if (n_params < 0) {
    fprintf(stderr, "bbbp_model_param_count: %s\n", bbbp_last_error());
    return 3;
  }

  /* parameters: one device allocation, a host table of device pointers in state_dict order */
  size_t param_bytes = 0, input_bytes = 0;
  float* host_params = (float*)read_file(argv[1], &param_bytes);
  unsigned char* host_inputs = (unsigned char*)read_file(argv[2], &input_bytes);
  if (!host_params || !host_inputs) {
    fprintf(stderr, "cannot read %s / %s\n", argv[1], argv[2]);
    return 1;
  }
  size_t total = 0;
  for (int i = 0; i < n_params; ++i) total += bbbp_model_param_numel(&desc, i);
  if (total * sizeof(float) != param_bytes) {
    fprintf(stderr, "%s holds %zu bytes, the model has %zu parameters + buffers\n", argv[1], param_bytes, total);
    return 1;
  }
  float* dev_params = NULL;
  CUDA_OK(cudaMalloc((void**)&dev_params, param_bytes));
  CUDA_OK(cudaMemcpy(dev_params, host_params, param_bytes, cudaMemcpyHostToDevice));
  const void** table = (const void**)malloc(sizeof(void*) * (size_t)n_params);
  size_t at = 0;
  for (int i = 0; i < n_params; ++i) {
    table[i] = dev_params + at;
    at += bbbp_model_param_numel(&desc, i);
  }

  const size_t rows = (size_t)desc.groups * desc.seq, F = (size_t)desc.fingerprint_size;
  const size_t fp_bytes = rows * F * sizeof(float);
  const size_t img_bytes = rows * 3 * 128 * 128 * (desc.image_is_u8 ? 1 : sizeof(float));
  if (fp_bytes + img_bytes != input_bytes) {
    fprintf(stderr, "%s holds %zu bytes, expected %zu\n", argv[2], input_bytes, fp_bytes + img_bytes);
    return 1;
  }
  void *dev_fp = NULL, *dev_img = NULL, *prepared = NULL, *workspace = NULL;
  float* dev_out = NULL;
  const size_t prepared_bytes = bbbp_model_prepared_bytes(&desc), workspace_bytes = bbbp_workspace_bytes(&desc);
  if (!prepared_bytes || !workspace_bytes) {
    fprintf(stderr, "bbbp_workspace_bytes: %s\n", bbbp_last_error());
    return 3;
  }
  CUDA_OK(cudaMalloc(&dev_fp, fp_bytes));
  CUDA_OK(cudaMalloc(&dev_img, img_bytes));
  CUDA_OK(cudaMalloc(&prepared, prepared_bytes));
  CUDA_OK(cudaMalloc(&workspace, workspace_bytes));
  CUDA_OK(cudaMalloc((void**)&dev_out, rows * sizeof(float)));
  CUDA_OK(cudaMemcpy(dev_fp, host_inputs, fp_bytes, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dev_img, host_inputs + fp_bytes, img_bytes, cudaMemcpyHostToDevice));

  cudaStream_t stream;
  CUDA_OK(cudaStreamCreate(&stream));
  BBBP_OK_OR_DIE(bbbp_model_prepare(&desc, table, prepared, prepared_bytes, stream));
  BBBP_OK_OR_DIE(bbbp_fwd(&desc, dev_fp, dev_img, table, prepared, dev_out, workspace, workspace_bytes, stream));

  /* the call is graph-capturable: capture it once and replay it (what a serving loop would do) */
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  BBBP_OK_OR_DIE(bbbp_fwd(&desc, dev_fp, dev_img, table, prepared, dev_out, workspace, workspace_bytes, stream));
  CUDA_OK(cudaStreamEndCapture(stream, &graph));
  CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
  cudaEvent_t t0, t1;
  CUDA_OK(cudaEventCreate(&t0));
  CUDA_OK(cudaEventCreate(&t1));
  CUDA_OK(cudaEventRecord(t0, stream));
  for (int i = 0; i < 10; ++i) CUDA_OK(cudaGraphLaunch(exec, stream));
  CUDA_OK(cudaEventRecord(t1, stream));
  CUDA_OK(cudaStreamSynchronize(stream));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, t0, t1));

  float* scores = (float*)malloc(rows * sizeof(float));
  CUDA_OK(cudaMemcpy(scores, dev_out, rows * sizeof(float), cudaMemcpyDeviceToHost));
  FILE* f = fopen(argv[3], "wb");
  if (!f || fwrite(scores, sizeof(float), rows, f) != rows) {
    fprintf(stderr, "cannot write %s\n", argv[3]);
    return 1;
  }
  fclose(f);
  printf("{\"molecules\": %zu, \"graph_replay_ms\": %.4f, \"molecules_per_s\": %.1f, \"launches\": %llu}\n", rows, ms / 10,
         rows / (ms / 10 * 1e-3), (unsigned long long)bbbp_launch_count());
  return 0;
}
