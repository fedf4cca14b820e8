/*
 * net_harness.c — runs the REFERENCE's whole device program (Runtime_Engine/cnn/device/src/cnn.cl, compiled
 * as plain C where it lies; nothing is copied) for ONE image through ALL layers of the network the tables
 * describe: input_reader, filter_reader, sequencer, retriever, the 16 PE kernels, relu, pool,
 * full_size_pool, pool_tail and feature_writer run CONCURRENTLY as coroutines, one stack each, connected by
 * the FIFO shim.  Test infrastructure only.
 *
 * This is what full_harness.c (layer 0 only) cannot do: layers > 0 read the on-chip feature cache that the
 * retriever fills from feature_writer / full_size_pool through NON-BLOCKING channel reads
 * (retriever.cl:328-329), and the ipool pseudo layers are fed from that cache (retriever.cl:285-302).
 * Scheduling model: a kernel runs until a blocking read finds its channel empty; a non-blocking read of an
 * empty channel first lets every other kernel run until all of them are parked ("the rest of the pipeline
 * is infinitely fast"), then takes what arrived.  Output tiles therefore reach the cache as early as the
 * data flow allows — never later than on the device, where the schedule (cycle.cl) leaves the pipeline
 * latency as slack — so every cache read sees the value the device would see.
 *
 * Built by oracle/build_ref.sh with -DRESNET50 / -DGOOGLENET / -DRESNET50_PRUNED into
 * oracle/_ref/libtf2ref_net_<net>.so (x86-64 only: the context switch is six pushes and a stack swap).
 */
#define TF2_CORO
#include "fifo_shim.h"

#include <sys/mman.h>

#include "cnn.cl"

#if !defined(__x86_64__)
#error "net_harness.c: x86-64 only"
#endif

/* ---- context switch ------------------------------------------------------------------------------ */
void tf2_ctx_switch(void** save_sp, void* new_sp);
__asm__(
    ".text\n.globl tf2_ctx_switch\n.type tf2_ctx_switch,@function\n"
    "tf2_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size tf2_ctx_switch, .-tf2_ctx_switch\n");

enum { ST_READY, ST_BLOCKED, ST_NB, ST_THROTTLED, ST_DONE };
#define MAX_CORO 40
#define STACK_BYTES ((size_t)768 << 20)
#define HIGH_WATER ((size_t)4 << 20) /* a producer is parked while its channel holds this much */
#define LOW_WATER ((size_t)1 << 20)

typedef struct {
  void* sp;
  int state;
  fifo_t* f;
  size_t need;
  void (*fn)(int);
  int arg;
  void* stack;
} coro_t;
static coro_t g_co[MAX_CORO];
static int g_nco = 0;
static coro_t* g_cur = NULL;
static void* g_sched_sp = NULL;
static unsigned long long g_pushes = 0, g_quiet_mark = ~0ull, g_switches = 0;

static void coro_yield(void) { g_switches++; tf2_ctx_switch(&g_cur->sp, g_sched_sp); }
static void coro_entry(void) {
  g_cur->fn(g_cur->arg);
  g_cur->state = ST_DONE;
  coro_yield();
  abort(); /* a finished kernel is never resumed */
}
static void coro_wait(fifo_t* f, size_t n) {
  while (f->head + n > f->tail) {
    g_cur->state = ST_BLOCKED;
    g_cur->f = f;
    g_cur->need = n;
    coro_yield();
  }
  g_cur->state = ST_READY;
}
static void coro_before_nb_read(fifo_t* f, size_t n) {
  if (g_pushes == g_quiet_mark) return; /* nothing was written since everybody else ran dry */
  g_cur->state = ST_NB;
  coro_yield();
  g_cur->state = ST_READY;
}
static void coro_after_push(fifo_t* f) {
  g_pushes++;
  if (f->tail - f->head > HIGH_WATER) {
    g_cur->state = ST_THROTTLED;
    g_cur->f = f;
    coro_yield();
    g_cur->state = ST_READY;
  }
}
static int coro_is_sink(const void* key) { /* the end of the PE daisy chain: pe_tail() only drains it (pe.cl:217-228) */
  return key == &pe_input_data_channel[N_VECTOR - 1] || key == &pe_input_filter_channel[N_VECTOR - 1] ||
         key == &pe_control_channel[N_VECTOR - 1];
}

/* ---- tap: every tile the retriever is sent for its cache ------------------------------------------ */
static unsigned char* g_tap = NULL;
static long long g_tap_cap = 0, g_tap_n = 0, g_tap_dropped = 0;
static void coro_tap(const void* key, const void* v, size_t n) {
  int which = key == (const void*)&retriever_input_channel ? 0 : key == (const void*)&end_pool_output_channel ? 1 : -1;
  if (which < 0 || !g_tap) return;
  if (g_tap_n >= g_tap_cap) { g_tap_dropped++; return; }
  unsigned char* rec = g_tap + g_tap_n * (8 + sizeof(PoolTailOutput));
  int hdr[2] = {which, 0};
  memcpy(rec, hdr, 8);
  memcpy(rec + 8, v, n);
  g_tap_n++;
}

static void coro_add(void (*fn)(int), int arg) {
  coro_t* c = &g_co[g_nco++];
  c->fn = fn;
  c->arg = arg;
  c->state = ST_READY;
  c->stack = mmap(NULL, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
  if (c->stack == MAP_FAILED) abort();
  void** top = (void**)((char*)c->stack + STACK_BYTES);
  top -= 8; /* [r15 r14 r13 r12 rbx rbp][return address = coro_entry][pad] */
  for (int i = 0; i < 6; i++) top[i] = NULL;
  top[6] = (void*)coro_entry;
  top[7] = NULL;
  c->sp = top;
}
static void coro_resume(coro_t* c) {
  g_cur = c;
  g_switches++;
  tf2_ctx_switch(&g_sched_sp, c->sp);
  g_cur = NULL;
}
static int coro_runnable(const coro_t* c) {
  switch (c->state) {
    case ST_READY: return 1;
    case ST_BLOCKED: return c->f->tail - c->f->head >= c->need;
    case ST_THROTTLED: return c->f->tail - c->f->head < LOW_WATER;
    default: return 0;
  }
}
static void coro_schedule(void) {
  for (;;) {
    int progress = 0;
    for (int i = 0; i < g_nco; i++)
      if (coro_runnable(&g_co[i])) { coro_resume(&g_co[i]); progress = 1; }
    if (progress) continue;
    g_quiet_mark = g_pushes; /* everybody is parked: non-blocking readers take what has arrived */
    for (int i = 0; i < g_nco; i++)
      if (g_co[i].state == ST_NB) { coro_resume(&g_co[i]); progress = 1; }
    if (progress) continue;
    for (int i = 0; i < g_nco && !progress; i++)
      if (g_co[i].state == ST_THROTTLED) { coro_resume(&g_co[i]); progress = 1; }
    if (!progress) return;
  }
}

/* ---- the kernels ----------------------------------------------------------------------------------- */
static const signed char* a_input;
static signed char* a_filter;
static BiasBnParam* a_bias_bn;
static int* a_idle;
static signed char* a_ddr;
static void k_input_reader(int x) { input_reader(1, (const real*)a_input); }
static void k_filter_reader(int x) { filter_reader(1, (real*)a_filter, a_bias_bn); }
static void k_sequencer(int x) { sequencer(1); }
static void k_retriever(int x) { retriever(1, a_idle); }
static void k_pe(int n) { PeFunction(n); }
static void k_relu(int x) { relu(1); }
static void k_pool(int x) { pool(1); }
static void k_full_size_pool(int x) { full_size_pool(1); }
static void k_pool_tail(int x) { pool_tail(1, (real*)a_ddr); }
static void k_feature_writer(int x) { feature_writer(1, (real*)a_ddr); }

int net_num_layer(void) { return NUM_LAYER; }
int net_tap_record_bytes(void) { return 8 + (int)sizeof(PoolTailOutput); }
int net_tap_data_offset(void) { return 8 + (int)((char*)&((PoolTailOutput*)0)->write_data - (char*)0); }
int net_tap_addr_offset(void) { return 8 + (int)((char*)&((PoolTailOutput*)0)->cache_write_addr - (char*)0); }
long long net_ddr_bytes(void) { return (long long)DDR_SIZE * NEXT_POWER_OF_2(W_VECTOR * NARROW_N_VECTOR) + 2ll * OUTPUT_OFFSET; }
long long net_output_offset(void) { return (long long)OUTPUT_OFFSET; }
/* per-layer facts the checker needs to cut the tap / the DDR image into tensors */
long long net_layer_info(int l, int which) {
  switch (which) {
    case 0: return kCacheWriteEnable[l];
    case 1: return kEndPoolEnable[l];
    case 2: return FEATURE_WRITER_CYCLE(l);
    case 3: return kDDRWriteEnable[l];
    case 4: return (long long)kDDRWriteBase[l] * NEXT_POWER_OF_2(W_VECTOR * NARROW_N_VECTOR);
    case 5: return kNStart[l];
    case 6: return kCacheWriteBase[l];
    case 7: return CONV_CYCLE(l);
    case 8: return kIpoolEnable[l];
    case 9: return kNvecEnd[l];
    case 10: return kPoolOutputHeight[l];
    case 11: return kPoolOutputWvecEnd[l];
    case 12: return kOutputChannels[l];
    case 13: return kPoolOutputWidth[l];
    default: return -1;
  }
}

/* input_buffer: int8 image in the InputConvert layout; gl_filter: FilterConvert output of the whole network;
 * bias_bn: BiasBnParam[NUM_CONVOLUTIONS * MAX_BIAS_SIZE]; idle: kSequencerIdleCycle as the host uploads it
 * (network.cpp:142-147); ddr: feature_ddr (net_ddr_bytes()); tap: records of net_tap_record_bytes().
 * stats[0..]: kernels finished, kernels left parked, context switches, tap records dropped, bytes left in FIFOs */
int net_run(const signed char* input_buffer, signed char* gl_filter, BiasBnParam* bias_bn, signed char* ddr,
            unsigned char* tap, long long tap_cap, long long* n_tap, long long* stats) {
  static int idle[NUM_CONVOLUTIONS];
  for (int l = 0; l < NUM_CONVOLUTIONS; l++) idle[l] = kSequencerIdleCycle[l];
  fifo_reset_all();
  a_input = input_buffer; a_filter = gl_filter; a_bias_bn = bias_bn; a_idle = idle; a_ddr = ddr;
  g_tap = tap; g_tap_cap = tap_cap; g_tap_n = 0; g_tap_dropped = 0;
  g_nco = 0; g_pushes = 0; g_quiet_mark = ~0ull; g_switches = 0;
  coro_add(k_input_reader, 0);
  coro_add(k_filter_reader, 0);
  coro_add(k_sequencer, 0);
  coro_add(k_retriever, 0);
  for (int n = 0; n < N_VECTOR; n++) coro_add(k_pe, n);
  coro_add(k_relu, 0);
  coro_add(k_pool, 0);
  coro_add(k_pool_tail, 0);
  coro_add(k_feature_writer, 0);
  coro_add(k_full_size_pool, 0);
  coro_schedule();
  long long done = 0, parked = 0, left = 0;
  for (int i = 0; i < g_nco; i++) {
    if (g_co[i].state == ST_DONE) done++; else { parked++; stats[8 + parked - 1] = i; }
    munmap(g_co[i].stack, STACK_BYTES);
  }
  for (int i = 0; i < g_nfifos; i++) left += (long long)(g_fifos[i].tail - g_fifos[i].head);
  stats[0] = done; stats[1] = parked; stats[2] = (long long)g_switches; stats[3] = g_tap_dropped; stats[4] = left;
  *n_tap = g_tap_n;
  g_tap = NULL;
  return 0;
}
