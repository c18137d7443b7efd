/*
 * libgg_pyhost.so -- host-side marshalling of the reference's input type for the graph build: a Python list of read
 * strings (KmerSsgnn.ssgnn_dataset, src/representations/kmer_ssgnn.py:56-57) -> one ASCII byte stream in page-locked
 * memory, the form the counting kernels read (include/gg_b200.h, K1).
 *
 * "\n".join(reads).encode() costs ~0.1 s for 1M reads of 150 bp in the interpreter -- several times the whole GPU
 * build.  Here the list is walked once under the GIL to collect (pointer, length) of every string's ASCII payload
 * (no copy: compact ASCII str objects expose it), and the bytes are then copied by native threads with the GIL
 * released, read range by read range, so that the copy of one range overlaps the upload and counting of the
 * previous one.  Loaded with ctypes.PyDLL (the calls need the GIL); it is the only part of the package that knows
 * about CPython, the CUDA library keeps a plain C ABI.
 *
 * Layout written: every read has the same length L  -> rows of L bytes back to back (fixed_len = L);
 *                 otherwise                         -> every read followed by one '\n' (fixed_len = 0).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GG_PY_API __attribute__((visibility("default")))

typedef struct {
  PyObject* seq;       /* PySequence_Fast result: keeps the items alive */
  const char** ptr;
  int64_t* off;        /* n + 1 stream offsets */
  int64_t n, n_bytes, fixed_len;
} gg_reads_t;

typedef struct {
  PyObject** items;
  gg_reads_t* h;
  int64_t lo, hi;
  int64_t bytes, first_len, bad;   /* results: payload bytes of the range, its common length (-2: mixed), first bad item */
} gg_scan_job_t;

/* Reads object headers only (type, kind flags, length, data pointer): safe without the GIL while the sequence -- which
 * the handle holds a reference to -- keeps every item alive. */
static void* gg_scan_worker(void* arg) {
  gg_scan_job_t* j = (gg_scan_job_t*)arg;
  int64_t bytes = 0, first_len = -1, bad = -1;
  for (int64_t i = j->lo; i < j->hi; ++i) {
    PyObject* it = j->items[i];
    if (!PyUnicode_Check(it) || !PyUnicode_IS_COMPACT_ASCII(it)) {
      if (bad < 0) bad = i;
      j->h->ptr[i] = NULL;
      j->h->off[i + 1] = 0;
      continue;
    }
    const int64_t len = (int64_t)PyUnicode_GET_LENGTH(it);
    j->h->ptr[i] = (const char*)PyUnicode_DATA(it);
    j->h->off[i + 1] = len;            /* lengths first; turned into offsets below */
    bytes += len;
    if (first_len == -1) first_len = len;
    else if (first_len != len) first_len = -2;
  }
  j->bytes = bytes;
  j->first_len = first_len;
  j->bad = bad;
  return NULL;
}

/* info: [0] reads, [1] stream bytes, [2] fixed_len (0: '\n' after every read), [3] index of the first item that is not
 * an ASCII str (-1: none).  Returns a handle (NULL on error, with a Python exception set or info[3] >= 0). */
GG_PY_API void* gg_py_reads_scan(PyObject* reads, int n_threads, int64_t* info) {
  info[0] = info[1] = info[2] = 0;
  info[3] = -1;
  PyObject* seq = PySequence_Fast(reads, "reads must be a list or tuple of str");
  if (!seq) return NULL;
  const int64_t n = (int64_t)PySequence_Fast_GET_SIZE(seq);
  gg_reads_t* h = (gg_reads_t*)calloc(1, sizeof(gg_reads_t));
  h->seq = seq;
  h->n = n;
  h->ptr = (const char**)malloc(sizeof(char*) * (size_t)(n > 0 ? n : 1));
  h->off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
  h->off[0] = 0;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  if (n < 65536) n_threads = 1;
  gg_scan_job_t jobs[64];
  pthread_t tid[64];
  PyObject** items = PySequence_Fast_ITEMS(seq);
  Py_BEGIN_ALLOW_THREADS
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].items = items;
    jobs[t].h = h;
    jobs[t].lo = n * t / n_threads;
    jobs[t].hi = n * (t + 1) / n_threads;
  }
  int started = 0;
  for (int t = 1; t < n_threads; ++t) {
    if (pthread_create(&tid[t], NULL, gg_scan_worker, &jobs[t]) != 0) break;
    started = t;
  }
  gg_scan_worker(&jobs[0]);
  for (int t = started + 1; t < n_threads; ++t) gg_scan_worker(&jobs[t]);
  for (int t = 1; t <= started; ++t) pthread_join(tid[t], NULL);
  Py_END_ALLOW_THREADS
  int64_t total = 0, first_len = -1, bad = -1;
  for (int t = 0; t < n_threads; ++t) {
    total += jobs[t].bytes;
    if (jobs[t].bad >= 0 && bad < 0) bad = jobs[t].bad;
    if (jobs[t].first_len == -1) continue;                 /* empty range */
    if (first_len == -1) first_len = jobs[t].first_len;
    else if (first_len != jobs[t].first_len) first_len = -2;
  }
  if (bad >= 0) {
    info[3] = bad;
    Py_DECREF(seq);
    free(h->ptr);
    free(h->off);
    free(h);
    return NULL;
  }
  const int fixed = n > 0 && first_len > 0;
  const int64_t sep = fixed ? 0 : 1;                        /* one separator after every read */
  for (int64_t i = 0; i < n; ++i) h->off[i + 1] += h->off[i] + sep;
  h->fixed_len = fixed ? first_len : 0;
  h->n_bytes = total + sep * n;
  info[0] = n;
  info[1] = h->n_bytes;
  info[2] = h->fixed_len;
  return h;
}

typedef struct {
  const gg_reads_t* h;
  uint8_t* out;
  int64_t lo, hi;
} gg_copy_job_t;

static void* gg_copy_worker(void* arg) {
  const gg_copy_job_t* j = (const gg_copy_job_t*)arg;
  const gg_reads_t* h = j->h;
  const int sep = h->fixed_len == 0;
  for (int64_t i = j->lo; i < j->hi; ++i) {
    const int64_t len = h->off[i + 1] - h->off[i] - sep;
    memcpy(j->out + h->off[i], h->ptr[i], (size_t)len);
    if (sep) j->out[h->off[i] + len] = '\n';
  }
  return NULL;
}

/* Copy reads [read_lo, read_hi) to out + their stream offsets (out = start of the WHOLE stream buffer, at least
 * info[1] bytes) on n_threads native threads, GIL released.  Returns 0, or -1 on a bad argument. */
GG_PY_API int gg_py_reads_copy(void* handle, int64_t read_lo, int64_t read_hi, uint8_t* out, int n_threads) {
  const gg_reads_t* h = (const gg_reads_t*)handle;
  if (!h || !out || read_lo < 0 || read_hi > h->n || read_lo > read_hi) return -1;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 64) n_threads = 64;
  const int64_t count = read_hi - read_lo;
  if (count == 0) return 0;
  if (count < 4096) n_threads = 1;
  gg_copy_job_t jobs[64];
  pthread_t tid[64];
  Py_BEGIN_ALLOW_THREADS
  for (int t = 0; t < n_threads; ++t) {
    jobs[t].h = h;
    jobs[t].out = out;
    jobs[t].lo = read_lo + count * t / n_threads;
    jobs[t].hi = read_lo + count * (t + 1) / n_threads;
  }
  int started = 0;
  for (int t = 1; t < n_threads; ++t) {
    if (pthread_create(&tid[t], NULL, gg_copy_worker, &jobs[t]) != 0) break;
    started = t;
  }
  gg_copy_worker(&jobs[0]);
  for (int t = started + 1; t < n_threads; ++t) gg_copy_worker(&jobs[t]);   /* threads that could not start */
  for (int t = 1; t <= started; ++t) pthread_join(tid[t], NULL);
  Py_END_ALLOW_THREADS
  return 0;
}

/* stream offset of read i (i == n: end of the stream) */
GG_PY_API int64_t gg_py_reads_offset(void* handle, int64_t i) {
  const gg_reads_t* h = (const gg_reads_t*)handle;
  return (!h || i < 0 || i > h->n) ? -1 : h->off[i];
}

GG_PY_API void gg_py_reads_free(void* handle) {
  gg_reads_t* h = (gg_reads_t*)handle;
  if (!h) return;
  Py_XDECREF(h->seq);
  free(h->ptr);
  free(h->off);
  free(h);
}
