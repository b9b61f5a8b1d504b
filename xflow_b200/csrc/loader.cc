// Host ingest (C ABI layer 4): the reference's block reader + libffm-style parser + feature hashing,
// re-implemented to emit a CSR batch (row_ptr u32, keys u64, labels u8) in pinned memory that the
// fused step consumes directly.
//
// Behaviour follows LoadData::load_minibatch_hash_data_fread (src/io/load_data_from_disk.cc:103-210):
//   * a block is at most block_bytes-1 bytes of text; when the buffer fills, it is cut after the
//     last '\n' and the remainder is carried into the next block (:108-124);
//   * row  = "<label>\t<tok> <tok> ...\n", label = ((float)atof(..) > 1e-7) (:131-135);
//   * tok  = "<fgid>:<fid>:<val>"; the feature key is std::hash<std::string>(<fid>) (:146-157,
//     io.h:46) — libstdc++'s MurmurHash64A with seed 0xc70f6907, restated in hash.h;
//     <val> is never read (the model treats x == 1, lr_worker.cc:132).
// Tokens must have all three fields (the reference's scan runs off the token otherwise).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/xflow_b200.h"
#include "hash.h"

void xf_set_error(const char* fmt, ...);

XF_DLL uint64_t xf_hash_bytes(const char* s, uint64_t len) { return xf_murmur64a(s, len); }

XF_DLL int xf_hash_decimal_ids(const uint64_t* ids, uint64_t n, uint64_t* out) {
  if ((!ids || !out) && n) return XF_ERR_ARG;
  for (uint64_t i = 0; i < n; ++i) {
    char buf[24];
    int len = 0;
    uint64_t v = ids[i];
    char tmp[24];
    do { tmp[len++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int j = 0; j < len; ++j) buf[j] = tmp[len - 1 - j];
    out[i] = xf_murmur64a(buf, (uint64_t)len);
  }
  return XF_OK;
}

struct xf_loader {
  FILE* fp = nullptr;
  uint64_t file_size = 0, file_pos = 0;  // regular files: reads go through pread on fileno(fp)
  bool regular = true;                   // FIFOs / pipes / stdin are read sequentially until EOF
  // two raw-text buffers, alternated by every block: the text of block i stays valid (e.g. as the source
  // of an asynchronous H2D copy, xf_trainer_ingest_begin) while block i+1 is being read
  char* buf = nullptr;                   // the current one
  char* buf_set[2] = {nullptr, nullptr};
  int buf_cur = 0;
  bool buf_pinned = false;
  size_t buf_size = 0, bmax = 0, btop = 0;
  // two output sets, alternated by every xf_loader_next: the arrays of block i stay valid (e.g. as
  // the source of an asynchronous H2D copy) while block i+1 is being parsed
  uint32_t* row_ptr_set[2] = {nullptr, nullptr};
  uint64_t* keys_set[2] = {nullptr, nullptr};
  uint8_t* labels_set[2] = {nullptr, nullptr};
  int cur = 1;
  uint32_t* row_ptr = nullptr;
  uint64_t* keys = nullptr;
  uint8_t* labels = nullptr;
  size_t max_rows = 0, max_tok = 0;
  bool pinned = false;
  uint32_t rows = 0, nnz = 0;
};

static void* xf_host_alloc(size_t bytes, bool* pinned) {
  void* p = nullptr;
  if (*pinned) {
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) return p;
    cudaGetLastError();
    *pinned = false;
  }
  return malloc(bytes);
}
static void xf_host_free(void* p, bool pinned) {
  if (!p) return;
  if (pinned) cudaFreeHost(p);
  else free(p);
}

XF_DLL int xf_loader_open(xf_loader** out, const char* path, uint64_t block_bytes) {
  if (!out || !path || block_bytes < 16) { xf_set_error("bad loader arguments"); return XF_ERR_ARG; }
  FILE* fp = fopen(path, "r");
  if (!fp) {
    xf_set_error("open file %s error!", path);  // io.h:33-36 (the reference exits here)
    return XF_ERR_IO;
  }
  xf_loader* l = new xf_loader;
  l->fp = fp;
  {
    struct stat sb;
    const bool ok = fstat(fileno(fp), &sb) == 0;
    l->regular = ok && S_ISREG(sb.st_mode);
    l->file_size = (ok && l->regular) ? (uint64_t)sb.st_size : 0;
  }
  l->buf_size = (size_t)block_bytes;
  {
    bool pin_text = true;  // the raw block is also what the device parser uploads (xf_loader_next_raw)
    l->buf_set[0] = (char*)xf_host_alloc(l->buf_size + 1, &pin_text);
    bool pin2 = pin_text;
    l->buf_set[1] = (char*)xf_host_alloc(l->buf_size + 1, &pin2);
    if (pin2 != pin_text && l->buf_set[1]) {  // keep both buffers of one kind
      xf_host_free(l->buf_set[1], pin2);
      l->buf_set[1] = nullptr;
    }
    l->buf_pinned = pin_text;
    l->buf = l->buf_set[0];
  }
  // shortest row "0\n" = 2 bytes, shortest token "a:b:c " ~ 4 bytes.  The CSR output sets of the HOST parser
  // are allocated by the first xf_loader_next: callers that only form raw blocks (device parser) never pay
  // for them.
  l->max_rows = l->buf_size / 2 + 2;
  l->max_tok = l->buf_size / 4 + 2;
  if (!l->buf_set[0] || !l->buf_set[1]) {
    xf_set_error("loader allocation failed");
    xf_loader_close(l);
    return XF_ERR_IO;
  }
  *out = l;
  return XF_OK;
}

XF_DLL int xf_loader_close(xf_loader* l) {
  if (!l) return XF_OK;
  if (l->fp) fclose(l->fp);
  xf_host_free(l->buf_set[0], l->buf_pinned);
  xf_host_free(l->buf_set[1], l->buf_pinned);
  for (int s = 0; s < 2; ++s) {
    xf_host_free(l->row_ptr_set[s], l->pinned);
    xf_host_free(l->keys_set[s], l->pinned);
    xf_host_free(l->labels_set[s], l->pinned);
  }
  delete l;
  return XF_OK;
}

// fread replacement: fills dst with up to n bytes from the current file position.  Large requests are
// split across a few threads (pread at disjoint offsets): one core copies out of the page cache at
// ~4 GB/s, which would otherwise cap file -> device throughput now that parsing runs on the GPU.
static size_t xf_read_block(xf_loader* l, char* dst, size_t n) {
  if (!l->regular) {
    // not seekable (FIFO, pipe, /dev/stdin): plain sequential reads until the buffer is full or EOF, like
    // the reference's fread (load_data_from_disk.cc:112)
    size_t done = 0;
    while (done < n) {
      const size_t r = fread(dst + done, 1, n - done, l->fp);
      if (r == 0) break;
      done += r;
    }
    return done;
  }
  const int fd = fileno(l->fp);
  const uint64_t remaining = l->file_size > l->file_pos ? l->file_size - l->file_pos : 0;
  if (n > remaining) n = (size_t)remaining;
  if (n == 0) return 0;
  const size_t kMinPerThread = (size_t)2 << 20;
  unsigned hw = std::thread::hardware_concurrency();
  size_t nthreads = std::min<size_t>(std::min<size_t>(16, hw ? hw : 1), n / kMinPerThread);
  if (nthreads < 1) nthreads = 1;
  std::vector<size_t> got(nthreads, 0);
  auto work = [&](size_t t) {
    const size_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
    size_t done = 0;
    while (lo + done < hi) {
      const ssize_t r = pread(fd, dst + lo + done, hi - lo - done, (off_t)(l->file_pos + lo + done));
      if (r <= 0) break;
      done += (size_t)r;
    }
    got[t] = done;
  };
  std::vector<std::thread> th;
  for (size_t t = 1; t < nthreads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  // bytes are valid up to the first short piece
  size_t total = 0;
  for (size_t t = 0; t < nthreads; ++t) {
    const size_t want = n * (t + 1) / nthreads - n * t / nthreads;
    total += got[t];
    if (got[t] < want) break;
  }
  l->file_pos += total;
  return total;
}

// block formation (load_data_from_disk.cc:108-124): returns the length of the parse region [0, end)
static size_t xf_loader_form_block(xf_loader* l) {
  // the tail carried over from the previous block moves to the front of the OTHER buffer
  char* prev = l->buf;
  l->buf_cur ^= 1;
  char* buf = l->buf = l->buf_set[l->buf_cur];
  if (l->bmax < l->btop) memcpy(buf, prev + l->bmax, l->btop - l->bmax);
  l->btop -= l->bmax;
  l->btop += xf_read_block(l, buf + l->btop, l->buf_size - 1 - l->btop);
  l->bmax = l->btop;
  size_t end;
  if (l->btop + 1 == l->buf_size) {
    while (l->bmax > 0 && buf[l->bmax - 1] != (char)EOF && buf[l->bmax - 1] != '\n') --l->bmax;
    if (l->bmax != 0) end = l->bmax - 1;
    else { l->bmax = l->btop; end = l->btop; }
  } else {
    end = l->bmax;
  }
  buf[end] = '\0';
  return end;
}

XF_DLL int xf_loader_next_raw(xf_loader* l, const char** text, uint64_t* len) {
  if (!l || !text || !len) return XF_ERR_ARG;
  const size_t end = xf_loader_form_block(l);
  *text = l->buf;
  *len = (uint64_t)end;
  return XF_OK;
}

static int xf_loader_alloc_sets(xf_loader* l) {
  if (l->row_ptr_set[0]) return XF_OK;
  int ndev = 0;
  bool pin = (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0);
  if (!pin) cudaGetLastError();
  l->pinned = pin;
  for (int s = 0; s < 2; ++s) {
    bool p1 = l->pinned, p2 = l->pinned, p3 = l->pinned;
    l->row_ptr_set[s] = (uint32_t*)xf_host_alloc((l->max_rows + 1) * 4, &p1);
    l->keys_set[s] = (uint64_t*)xf_host_alloc(l->max_tok * 8, &p2);
    l->labels_set[s] = (uint8_t*)xf_host_alloc(l->max_rows, &p3);
    if (!l->row_ptr_set[s] || !l->keys_set[s] || !l->labels_set[s] || !(p1 == p2 && p2 == p3) || (s == 1 && p1 != l->pinned)) {
      xf_set_error("loader allocation failed");
      return XF_ERR_IO;
    }
    l->pinned = p1;
  }
  return XF_OK;
}

// back to the first byte of the file (regular files only): the next block is the first block again.  Lets a
// caller run every epoch on one loader instead of re-allocating page-locked buffers per epoch.
XF_DLL int xf_loader_rewind(xf_loader* l) {
  if (!l) return XF_ERR_ARG;
  if (!l->regular) { xf_set_error("loader: cannot rewind a stream"); return XF_ERR_IO; }
  l->file_pos = 0;
  l->bmax = l->btop = 0;
  return XF_OK;
}

XF_DLL int xf_loader_next(xf_loader* l, uint32_t* rows_out, uint32_t* nnz_out) {
  if (!l || !rows_out || !nnz_out) return XF_ERR_ARG;
  if (xf_loader_alloc_sets(l) != XF_OK) return XF_ERR_IO;
  l->cur ^= 1;
  l->row_ptr = l->row_ptr_set[l->cur];
  l->keys = l->keys_set[l->cur];
  l->labels = l->labels_set[l->cur];
  const size_t end = xf_loader_form_block(l);
  char* buf = l->buf;

  // --- parse
  uint32_t rows = 0, nnz = 0;
  const char* p = buf;
  const char* const stop = buf + end;
  l->row_ptr[0] = 0;
  while (p < stop && *p != '\0') {
    const char* tab = (const char*)memchr(p, '\t', (size_t)(stop - p));
    if (!tab) break;
    // label: (float)atof(text) > 1e-7.  Fast path for the ubiquitous "0" / "1".
    uint8_t y;
    if (tab - p == 1 && (*p == '0' || *p == '1')) {
      y = (uint8_t)(*p - '0');
    } else {
      char tmp[64];
      size_t ll = (size_t)(tab - p) < sizeof(tmp) - 1 ? (size_t)(tab - p) : sizeof(tmp) - 1;
      memcpy(tmp, p, ll);
      tmp[ll] = '\0';
      float yf = (float)atof(tmp);
      y = (yf > 0.0000001) ? 1 : 0;
    }
    if (rows >= l->max_rows) { xf_set_error("loader: row capacity exceeded"); return XF_ERR_IO; }
    l->labels[rows] = y;
    p = tab + 1;
    const char* eol = (const char*)memchr(p, '\n', (size_t)(stop - p));
    const char* const line_end = eol ? eol : stop;
    while (p < line_end) {
      // token = [p, q): up to the next ' ' or end of line
      const char* c1 = nullptr;
      const char* c2 = nullptr;
      const char* q = p;
      for (; q < line_end && *q != ' '; ++q) {
        if (*q == ':') {
          if (!c1) c1 = q;
          else if (!c2) c2 = q;
        }
      }
      // a token that starts with '\r' is the CR of a CRLF row without features, not a feature (the device
      // parser in ingest.cu applies the same rule)
      if (q > p && *p != '\r') {
        if (!c1 || !c2) { xf_set_error("loader: token without three ':'-separated fields"); return XF_ERR_IO; }
        if (nnz >= l->max_tok) { xf_set_error("loader: token capacity exceeded"); return XF_ERR_IO; }
        l->keys[nnz++] = xf_murmur64a(c1 + 1, (uint64_t)(c2 - c1 - 1));
      }
      p = (q < line_end) ? q + 1 : line_end;
    }
    l->row_ptr[++rows] = nnz;
    p = eol ? eol + 1 : stop;
  }
  l->rows = rows;
  l->nnz = nnz;
  *rows_out = rows;
  *nnz_out = nnz;
  return XF_OK;
}

XF_DLL int xf_loader_batch(xf_loader* l, const uint32_t** row_ptr, const uint64_t** keys, const uint8_t** labels) {
  if (!l) return XF_ERR_ARG;
  if (row_ptr) *row_ptr = l->row_ptr;
  if (keys) *keys = l->keys;
  if (labels) *labels = l->labels;
  return XF_OK;
}
