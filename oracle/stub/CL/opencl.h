/* Minimal stand-in for the Intel FPGA OpenCL SDK header so that the reference's HOST loader
 * sources (model_loader.cpp, quantization.cpp, input_loader.cpp) compile unmodified here.
 * Types only — nothing OpenCL is ever called by those three files.  Test infrastructure. */
#ifndef TF2B_STUB_OPENCL_H
#define TF2B_STUB_OPENCL_H
#include <stddef.h>
#include <stdint.h>
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_uint cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_uint cl_device_info;
typedef cl_uint cl_platform_info;
typedef struct _cl_platform_id* cl_platform_id;
typedef struct _cl_device_id* cl_device_id;
typedef struct _cl_context* cl_context;
typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_mem* cl_mem;
typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel;
typedef struct _cl_event* cl_event;
#define CL_SUCCESS 0
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_CALLBACK
#define CL_DEVICE_TYPE_ALL 0xFFFFFFFF
#endif
