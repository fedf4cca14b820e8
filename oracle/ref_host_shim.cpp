// Link shim for the compiled-reference host loaders (oracle/_ref/libtf2ref_host_<net>.so).
// Provides the two aocl_utils symbols nothing in the loaders really needs plus extern "C" entry
// points so ctypes can call the reference's own C++ functions.  Test infrastructure only.
#include <cstdlib>
#include "includes.h"
#include "network_helper.h"

namespace aocl_utils {
void* alignedMalloc(size_t size) { void* p = nullptr; if (posix_memalign(&p, 64, size)) return nullptr; return p; }
void alignedFree(void* p) { free(p); }
}  // namespace aocl_utils

void filter_trans(real* filter_input, real* Transform_input);
char Get_real(float data, char expand);
void feature_trans(float* Feas, float* finals_feas);

extern "C" {
int ref_num_layer() { return NUM_LAYER; }
int ref_max_out_channel() { return MAX_OUT_CHANNEL; }
int ref_max_bias_size() { return MAX_BIAS_SIZE; }
long long ref_filter_layer_stride() { return (long long)MAX_FILTER_SIZE * NEXT_POWER_OF_2(FW_VECTOR * C_VECTOR); }
int ref_num_q_layers() { return NUM_Q_LAYERS; }
char ref_get_real(float w, char expand) { return Get_real(w, expand); }
void ref_filter_trans(char* in49, char* out81) { filter_trans(in49, out81); }
void ref_feature_trans(float* in, float* out) { feature_trans(in, out); }
void ref_quantization(char* q, char* file_name) { Quantization(q, nullptr, file_name); }
void ref_load_model(char* filename, char* filter_raw, BiasBnParam* bias_bn, char* q) { LoadModel(filename, filter_raw, bias_bn, q); }
void ref_input_convert(float* input_raw, float* input, int num_images) { InputConvert(input_raw, input, num_images); }
long long ref_input_device_size() {
  return (long long)CEIL(kInputChannels[0], C_VECTOR) * kInputHeight[0] * CEIL(kInputWidth[0], W_VECTOR) * NEXT_POWER_OF_2(W_VECTOR * C_VECTOR);
}
long long ref_filter_device_size() { return (long long)NUM_CONVOLUTIONS * MAX_FILTER_SIZE * NEXT_POWER_OF_2(FW_VECTOR * C_VECTOR); }
void ref_filter_convert(char* scratch, char* filter_raw, char* filter_real) { FilterConvert(scratch, filter_raw, filter_real); }
// network_helper.cpp:18-207 (both read the tiled feature_ddr image; Verify writes Lastconv<n>.dat in the cwd)
void ref_evaluation(int n, char* q, char* output, int* top_labels) { Evaluation(n, q, (real*)output, top_labels); }
void ref_verify(int n, char* file_name, char* q, char* output) { Verify(n, file_name, q, (real*)output); }
long long ref_output_offset() { return (long long)OUTPUT_OFFSET; }
long long ref_last_ddr_write_base() { return (long long)kDDRWriteBase[NUM_LAYER - 1] * NEXT_POWER_OF_2(W_VECTOR * NARROW_N_VECTOR); }
void ref_load_input_image(char* image_name, float* input_raw, float* raw_images) { LoadInputImage(image_name, input_raw, raw_images, 0); }
}
