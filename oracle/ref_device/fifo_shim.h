/*
 * fifo_shim.h — lets the REFERENCE's OpenCL device sources (Runtime_Engine/cnn/device/src/*.cl)
 * compile and run as plain C where they lie (nothing is copied).  Test infrastructure only.
 *
 * `channel` objects become statics; read/write_channel_altera become operations on unbounded FIFOs
 * keyed by the channel's address; reading an empty FIFO leaves the kernel through longjmp (the
 * kernel has consumed all of its input).  Kernels are then ordinary C functions that the harness
 * runs one after another in dataflow order.
 */
#ifndef TF2_FIFO_SHIM_H
#define TF2_FIFO_SHIM_H
#include <setjmp.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;

#define OPENCL
#define DISABLE_INFINITE_LOOPS
#define DISABLE_AUTORUN_KERNELS
#define constant static const
#define kernel
#define global
#define restrict __restrict__
#define channel static
static inline int max(int a, int b) { return a > b ? a : b; }

typedef struct {
  const void* key;
  unsigned char* buf;
  size_t head, tail, cap;
} fifo_t;
static fifo_t g_fifos[256];
static int g_nfifos = 0;
static jmp_buf g_exit;

static fifo_t* fifo_of(const void* key) {
  for (int i = 0; i < g_nfifos; i++)
    if (g_fifos[i].key == key) return &g_fifos[i];
  fifo_t* f = &g_fifos[g_nfifos++];
  f->key = key;
  f->buf = NULL;
  f->head = f->tail = f->cap = 0;
  return f;
}
/* optional cap on the number of items one channel accepts: the writer leaves through longjmp when it
 * is reached (used to stop the sequencer after the first layer's schedule) */
static const void* g_limit_key = NULL;
static size_t g_limit_bytes = 0;
#ifdef TF2_CORO
/* coroutine mode (net_harness.c): every kernel runs on its own stack, an empty blocking read parks the
 * kernel until its producer has written, a non-blocking read first lets every other kernel run dry */
static void coro_wait(fifo_t* f, size_t n);
static void coro_before_nb_read(fifo_t* f, size_t n);
static void coro_after_push(fifo_t* f);
static int coro_is_sink(const void* key);
static void coro_tap(const void* key, const void* v, size_t n);
#endif
static void fifo_push(const void* key, const void* v, size_t n) {
#ifdef TF2_CORO
  coro_tap(key, v, n);
  if (coro_is_sink(key)) return;
#endif
  fifo_t* f = fifo_of(key);
  if (key == g_limit_key && f->tail + n > g_limit_bytes) longjmp(g_exit, 2);
#ifdef TF2_CORO
  if (f->head == f->tail) f->head = f->tail = 0; /* consumed space is reused: whole networks stream GBs */
  else if (f->head > (1u << 20) && f->head > f->cap / 2) {
    memmove(f->buf, f->buf + f->head, f->tail - f->head);
    f->tail -= f->head;
    f->head = 0;
  }
#endif
  if (f->tail + n > f->cap) {
    f->cap = f->cap ? f->cap * 2 : (1u << 20);
    while (f->tail + n > f->cap) f->cap *= 2;
    f->buf = (unsigned char*)realloc(f->buf, f->cap);
  }
  memcpy(f->buf + f->tail, v, n);
  f->tail += n;
#ifdef TF2_CORO
  coro_after_push(f);
#endif
}
static void fifo_pop(const void* key, void* v, size_t n) {
  fifo_t* f = fifo_of(key);
#ifdef TF2_CORO
  if (f->head + n > f->tail) coro_wait(f, n);
#else
  if (f->head + n > f->tail) longjmp(g_exit, 1);
#endif
  memcpy(v, f->buf + f->head, n);
  f->head += n;
}
static size_t fifo_count(const void* key, size_t n) {
  fifo_t* f = fifo_of(key);
  return (f->tail - f->head) / n;
}
static void fifo_reset_all(void) {
  for (int i = 0; i < g_nfifos; i++) g_fifos[i].head = g_fifos[i].tail = 0;
}

#define read_channel_altera(c) ({ __typeof__(c) v__; fifo_pop(&(c), &v__, sizeof v__); v__; })
#define write_channel_altera(c, v) do { __typeof__(c) t__ = (v); fifo_push(&(c), &t__, sizeof t__); } while (0)
#ifdef TF2_CORO
static int fifo_try_pop(const void* key, void* v, size_t n) {
  fifo_t* f = fifo_of(key);
  if (f->head + n > f->tail) coro_before_nb_read(f, n);
  if (f->head + n > f->tail) { memset(v, 0, n); return 0; }
  memcpy(v, f->buf + f->head, n);
  f->head += n;
  return 1;
}
#define read_channel_nb_altera(c, valid) ({ __typeof__(c) v__; *(valid) = fifo_try_pop(&(c), &v__, sizeof v__) != 0; v__; })
#else
#define read_channel_nb_altera(c, valid) ({ __typeof__(c) v__; memset(&v__, 0, sizeof v__); *(valid) = false; v__; })
#endif
#endif
