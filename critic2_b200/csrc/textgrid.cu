// textgrid.cu -- formatted-text grid readers on sm_100a: the numeric block of a Gaussian cube file
// (read_cube, critic2 src/grid3mod@proc.f90:512-568: list-directed read of (((f(i,j,k),k=1,n3),j=1,n2),i=1,n1))
// and of a VASP CHGCAR/CHG/ELFCAR file (read_vasp, :842-913: (((f(i,j,k),i=1,n1),j=1,n2),k=1,n3), optionally divided
// by the cell volume).  SURVEY.md 8(f) rank 1: once the QTAIM kernels take milliseconds, parsing 10^8..10^9 formatted
// numbers on one CPU thread is what the user waits for.
//
// The caller (the unchanged Fortran reader) still parses the few header lines; it hands over the bytes that follow
// the header and the number of values to read.  On the device:
//   k_tok_count  one pass over the bytes: token starts (non-separator after a separator) per 4 KB chunk;
//   k_tok_scan   exclusive scan of the chunk counts;
//   k_tok_parse  every chunk is staged in shared memory, its token starts are compacted, and each thread converts
//                whole tokens: sign, up to 19 significant digits as a 64-bit integer, decimal exponent.
// Conversion is CORRECTLY ROUNDED (what a list-directed READ / strtod gives):
//   * m < 2^53 and |e10| <= 22: one IEEE multiplication or division by an exact power of ten (Clinger's fast path);
//   * otherwise m * 10^e10 in double-double arithmetic (106-bit powers of ten from pow10_dd.h, error < 2^-98); the
//     high word is the correctly rounded result unless the low word is within that error of a rounding boundary --
//     those tokens, subnormal results, overflows and mantissas longer than 19 digits are listed and converted by
//     strtod on the host (never seen on cube / CHGCAR data; the tests force them).
// Separators: blank, tab, newline, carriage return, comma.  Exponent letters E, D, Q (either case) or a bare sign
// ("1.5-03").  The r*c repeat form of list-directed input is not supported (neither format writes it).
#include "common.cuh"

#include <cstdlib>
#include <string>

#include "pow10_dd.h"

namespace {

constexpr int TCHUNK = 4096;   // bytes per block
constexpr int TTAIL = 64;      // bytes of the next chunk staged too (longest token that may straddle)
constexpr int TTHREADS = 256;  // 16 bytes per thread

__device__ __forceinline__ bool is_sep(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == ','; }

// token starts among the 16 bytes of a thread: bit q set if byte q is a non-separator preceded by a separator
__device__ __forceinline__ unsigned start_mask16(const unsigned char* b, unsigned char prev) {
  unsigned m = 0;
  bool psep = is_sep(prev);
#pragma unroll
  for (int q = 0; q < 16; q++) {
    const bool s = is_sep(b[q]);
    if (!s && psep) m |= 1u << q;
    psep = s;
  }
  return m;
}

__global__ void __launch_bounds__(TTHREADS) k_tok_count(const unsigned char* __restrict__ text, size_t nbytes, int* __restrict__ cnt) {
  __shared__ int s_w[TTHREADS / 32];
  const size_t base = (size_t)blockIdx.x * TCHUNK + (size_t)threadIdx.x * 16;
  // the buffer is padded with blanks to a multiple of TCHUNK + TTAIL: no bounds checks on the loads
  const uint4 v = *reinterpret_cast<const uint4*>(text + base);
  const unsigned char prev = base == 0 ? (unsigned char)' ' : text[base - 1];
  int c = __popc(start_mask16(reinterpret_cast<const unsigned char*>(&v), prev));
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < TTHREADS / 32; w++) t += s_w[w];
    cnt[blockIdx.x] = t;
  }
}

// exclusive scan of the chunk counts (64-bit totals), one block
__global__ void __launch_bounds__(1024) k_tok_scan(int nchunk, const int* __restrict__ cnt, long long* __restrict__ off, long long* __restrict__ total) {
  __shared__ long long s_w[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nchunk; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const long long v = i < nchunk ? cnt[i] : 0;
    long long incl = v;
    for (int d = 1; d < 32; d <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    long long wbase = 0;
    for (int w = 0; w < wid; w++) wbase += s_w[w];
    const long long carry = s_carry;
    if (i < nchunk) off[i] = carry + wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

struct ParseArgs {
  const unsigned char* text;
  size_t nbytes;
  const long long* off;      // first token index of every chunk
  long long nvalues;
  int n1, n2, n3, order;     // order 0: i fastest in the file (CHGCAR), 1: k fastest (cube)
  double divisor;            // every value is divided by this (1 = no scaling)
  double* out;
  // tokens to be converted on the host: (token index, byte offset)
  long long* fb_tok; unsigned long long* fb_pos; int* nfb; int fbcap;
  int* err;                  // 0 ok, 1 bad character, 2 token too long
  unsigned long long* errpos;
  unsigned long long* consumed;  // byte offset just past the last value read
};

__device__ __forceinline__ void two_prod(double a, double b, double& p, double& e) { p = a * b; e = fma(a, b, -p); }
__device__ __forceinline__ void fast_two_sum(double a, double b, double& s, double& e) { s = a + b; e = b - (s - a); }

// m * 10^e10 correctly rounded, or false if the decision needs more precision / the result is not a normal number
__device__ __forceinline__ bool dec_to_double(unsigned long long m, int e10, double& v) {
  if (m == 0) { v = 0.0; return true; }
  if (m < (1ull << 53)) {
    const double md = (double)(long long)m;
    if (e10 == 0) { v = md; return true; }
    if (e10 > 0 && e10 <= 22) { v = md * c2g_pow10_hi[e10]; return true; }   // both exact: one rounding
    if (e10 < 0 && e10 >= -22) { v = md / c2g_pow10_hi[-e10]; return true; }
  }
  const int ae = e10 < 0 ? -e10 : e10;
  if (ae > C2G_POW10_MAX) return false;
  // m as a double-double (exact: m < 2^64)
  double mh = (double)m;  // round to nearest
  const double ml = (double)(long long)(m - (unsigned long long)mh);
  const double ph = c2g_pow10_hi[ae], pl = c2g_pow10_lo[ae];
  double rh, rl;
  if (e10 >= 0) {
    double p, e;
    two_prod(mh, ph, p, e);
    e += mh * pl + ml * ph;
    fast_two_sum(p, e, rh, rl);
  } else {
    const double q1 = mh / ph;
    double p, e;
    two_prod(q1, ph, p, e);
    const double r = (((mh - p) - e) + ml) - q1 * pl;
    const double q2 = r / ph;
    fast_two_sum(q1, q2, rh, rl);
  }
  if (!(rh > 1e-290 && rh < 1e300)) return false;  // keep clear of subnormals / overflow: the host decides
  // ulp of rh and the distance of the low word from a rounding boundary
  const int ex = (__double2hiint(rh) >> 20) & 0x7ff;
  const double ulp = __hiloint2double((ex - 52) << 20, 0);
  const double arl = fabs(rl), tol = rh * 3.2e-30;  // 2^-98
  if (fabs(arl - 0.5 * ulp) <= tol) return false;
  const bool pow2 = (__double2hiint(rh) & 0xfffff) == 0 && __double2loint(rh) == 0;
  if (pow2 && fabs(arl - 0.25 * ulp) <= tol) return false;
  v = rh;
  return true;
}

__global__ void __launch_bounds__(TTHREADS) k_tok_parse(const __grid_constant__ ParseArgs A) {
  __shared__ __align__(16) unsigned char s_b[TCHUNK + TTAIL];
  __shared__ unsigned short s_tok[TCHUNK / 2 + 1];
  __shared__ int s_w[TTHREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const size_t cbase = (size_t)blockIdx.x * TCHUNK;
  const long long tok0 = A.off[blockIdx.x];
  if (tok0 >= A.nvalues) return;  // everything asked for lies before this chunk
  // stage the chunk and the head of the next one
  reinterpret_cast<uint4*>(s_b)[tid] = *reinterpret_cast<const uint4*>(A.text + cbase + (size_t)tid * 16);
  if (tid < TTAIL / 16) reinterpret_cast<uint4*>(s_b + TCHUNK)[tid] = *reinterpret_cast<const uint4*>(A.text + cbase + TCHUNK + (size_t)tid * 16);
  const size_t gpos = cbase + (size_t)tid * 16;
  const unsigned char prev = gpos == 0 ? (unsigned char)' ' : A.text[gpos - 1];
  __syncthreads();
  const unsigned m = start_mask16(s_b + tid * 16, prev);
  // exclusive scan of the per-thread start counts
  const int c = __popc(m);
  int incl = c;
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_w[wid] = incl;
  __syncthreads();
  int wbase = 0, ntok = 0;
  for (int w = 0; w < TTHREADS / 32; w++) {
    if (w < wid) wbase += s_w[w];
    ntok += s_w[w];
  }
  {
    int pos = wbase + incl - c;
    unsigned mm = m;
    while (mm) {
      const int q = __ffs(mm) - 1;
      mm &= mm - 1;
      s_tok[pos++] = (unsigned short)(tid * 16 + q);
    }
  }
  __syncthreads();
  for (int t = tid; t < ntok; t += TTHREADS) {
    const long long g = tok0 + t;
    if (g >= A.nvalues) break;
    int p = s_tok[t];
    const int pstart = p;
    // ---- sign, digits, fraction, exponent ----
    bool neg = false, bad = false;
    unsigned char ch = s_b[p];
    if (ch == '+' || ch == '-') { neg = ch == '-'; ch = s_b[++p]; }
    unsigned long long mant = 0;
    int nd = 0, dropped = 0, fracdigits = 0, ndig_total = 0;
    bool seen_dot = false, inexact = false;
    for (;; ch = s_b[++p]) {
      if (p >= TCHUNK + TTAIL - 1) break;
      if (ch >= '0' && ch <= '9') {
        ndig_total++;
        if (nd < 19) {
          mant = mant * 10ull + (unsigned)(ch - '0');
          if (mant != 0) nd++;
          if (seen_dot) fracdigits++;
        } else {  // more than 19 significant digits: the tail decides at most the rounding -> host
          if (ch != '0') inexact = true;
          if (!seen_dot) dropped++;
        }
      } else if (ch == '.' && !seen_dot) {
        seen_dot = true;
      } else {
        break;
      }
    }
    int e10 = 0;
    if (ch == 'e' || ch == 'E' || ch == 'd' || ch == 'D' || ch == 'q' || ch == 'Q' || ((ch == '+' || ch == '-') && ndig_total > 0)) {
      if (!(ch == '+' || ch == '-')) ch = s_b[++p];
      bool eneg = false;
      if (ch == '+' || ch == '-') { eneg = ch == '-'; ch = s_b[++p]; }
      int ne = 0;
      while (ch >= '0' && ch <= '9' && p < TCHUNK + TTAIL - 1) {
        if (e10 < 100000) e10 = e10 * 10 + (ch - '0');
        ne++;
        ch = s_b[++p];
      }
      if (ne == 0) bad = true;
      if (eneg) e10 = -e10;
    }
    if (ndig_total == 0) bad = true;
    if (!is_sep(ch)) bad = true;
    if (p >= TCHUNK + TTAIL - 1) {  // ran into the end of the staged bytes
      atomicMax(A.err, 2);
      atomicMin(A.errpos, (unsigned long long)(cbase + pstart));
      continue;
    }
    if (bad) {
      atomicMax(A.err, 1);
      atomicMin(A.errpos, (unsigned long long)(cbase + pstart));
      continue;
    }
    if (g == A.nvalues - 1) *A.consumed = (unsigned long long)(cbase + p);
    // ---- value ----
    double v = 0.0;
    const bool ok = !inexact && dec_to_double(mant, e10 - fracdigits + dropped, v);
    // destination in Fortran order
    size_t idx;
    if (A.order == 0) idx = (size_t)g;
    else {
      const long long k = g % A.n3, r = g / A.n3;
      const long long j = r % A.n2, i = r / A.n2;
      idx = (size_t)i + (size_t)A.n1 * ((size_t)j + (size_t)A.n2 * (size_t)k);
    }
    if (ok) {
      if (neg) v = -v;
      if (A.divisor != 1.0) v = v / A.divisor;
      A.out[idx] = v;
    } else {
      const int slot = atomicAdd(A.nfb, 1);
      if (slot < A.fbcap) { A.fb_tok[slot] = (long long)idx; A.fb_pos[slot] = (unsigned long long)(cbase + pstart); }
    }
  }
}

__global__ void k_patch(int n, const long long* __restrict__ idx, const double* __restrict__ val, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[idx[t]] = val[t];
}

}  // namespace

// Parses the first n1*n2*n3 numbers of `text` (host memory, nbytes bytes) into a new resident grid.
extern "C" int c2g_grid_parse_text(c2g_context* ctx, const char* text, size_t nbytes, const int n[3], int order, double divisor,
                                   int* handle, size_t* consumed, long long* nhost) {
  if (!ctx) return C2G_ERR_ARG;
  if (!text || !n || !handle) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: null argument");
  if (order != C2G_TEXT_ORDER_I_FASTEST && order != C2G_TEXT_ORDER_K_FASTEST) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: bad order %d", order);
  if (divisor == 0.0) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: zero divisor");
  int rc = c2g_grid_alloc(ctx, n, handle);
  if (rc != C2G_OK) return rc;
  struct HG { c2g_context* c; int h; bool ok = false; ~HG() { if (!ok) c2g_grid_free(c, h); } } hg{ctx, *handle};
  const long long nvalues = ctx->grids[*handle].nn;
  cudaStream_t st = ctx->stream;
  const size_t nchunk = (nbytes + TCHUNK - 1) / TCHUNK;
  if (nchunk == 0) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: empty text");
  if (nchunk > 0x7fffffffull) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: text too large");
  const size_t padded = nchunk * TCHUNK + TTAIL + 16;
  DevBuf b_text, b_cnt, b_off, b_ctl, b_fbt, b_fbp;
  C2G_CUDA(ctx, b_text.alloc(ctx, padded));
  C2G_CUDA(ctx, cudaMemsetAsync((char*)b_text.p + nbytes, ' ', padded - nbytes, st));
  ctx->prof_begin("text_h2d");
  C2G_CUDA(ctx, cudaMemcpyAsync(b_text.p, text, nbytes, cudaMemcpyHostToDevice, st));
  ctx->prof_end(0);
  C2G_CUDA(ctx, b_cnt.alloc(ctx, sizeof(int) * nchunk));
  C2G_CUDA(ctx, b_off.alloc(ctx, sizeof(long long) * nchunk));
  C2G_CUDA(ctx, b_ctl.alloc(ctx, 64));
  const int fbcap = 1 << 20;
  C2G_CUDA(ctx, b_fbt.alloc(ctx, sizeof(long long) * fbcap));
  C2G_CUDA(ctx, b_fbp.alloc(ctx, sizeof(unsigned long long) * fbcap));
  // ctl (8-byte slots): [0] total tokens, [1] errpos, [2] consumed, [3] err (int) | nfb (int)
  unsigned long long hctl[8] = {0, ~0ull, 0, 0, 0, 0, 0, 0};
  C2G_CUDA(ctx, cudaMemcpyAsync(b_ctl.p, hctl, 64, cudaMemcpyHostToDevice, st));
  unsigned long long* ctl = b_ctl.as<unsigned long long>();
  ctx->prof_begin("text_count");
  k_tok_count<<<(unsigned)nchunk, TTHREADS, 0, st>>>((const unsigned char*)b_text.p, nbytes, b_cnt.as<int>());
  k_tok_scan<<<1, 1024, 0, st>>>((int)nchunk, b_cnt.as<int>(), b_off.as<long long>(), (long long*)ctl);
  ctx->prof_end(2);
  C2G_KERNEL_CHECK(ctx);
  ParseArgs A;
  A.text = (const unsigned char*)b_text.p; A.nbytes = nbytes; A.off = b_off.as<long long>(); A.nvalues = nvalues;
  A.n1 = n[0]; A.n2 = n[1]; A.n3 = n[2]; A.order = order; A.divisor = divisor; A.out = ctx->grids[*handle].d;
  A.fb_tok = b_fbt.as<long long>(); A.fb_pos = b_fbp.as<unsigned long long>(); A.fbcap = fbcap;
  A.errpos = ctl + 1; A.consumed = ctl + 2; A.err = (int*)(ctl + 3); A.nfb = (int*)(ctl + 3) + 1;
  ctx->prof_begin("text_parse");
  k_tok_parse<<<(unsigned)nchunk, TTHREADS, 0, st>>>(A);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  C2G_CUDA(ctx, cudaMemcpyAsync(hctl, ctl, 64, cudaMemcpyDeviceToHost, st));
  C2G_CUDA(ctx, cudaStreamSynchronize(st));
  const long long ntok = (long long)hctl[0];
  const int err = (int)(hctl[3] & 0xffffffffull), nfb = (int)(hctl[3] >> 32);
  if (err == 1) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: not a number at byte %llu", hctl[1]);
  if (err == 2) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: token longer than %d bytes at byte %llu", TTAIL - 2, hctl[1]);
  if (ntok < nvalues) return ctx->fail(C2G_ERR_ARG, "c2g_grid_parse_text: %lld values expected, %lld found", nvalues, ntok);
  if (nfb > fbcap) return ctx->fail(C2G_ERR_OVERFLOW, "c2g_grid_parse_text: %d values need host conversion (limit %d)", nfb, fbcap);
  if (nfb > 0) {  // rare tokens: exact conversion with strtod on the host
    std::vector<long long> ftok(nfb);
    std::vector<unsigned long long> fpos(nfb);
    C2G_CUDA(ctx, cudaMemcpyAsync(ftok.data(), b_fbt.p, sizeof(long long) * nfb, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaMemcpyAsync(fpos.data(), b_fbp.p, sizeof(unsigned long long) * nfb, cudaMemcpyDeviceToHost, st));
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
    std::vector<double> fval(nfb);
    for (int q = 0; q < nfb; q++) {
      std::string tk;
      for (size_t p = fpos[q]; p < nbytes && tk.size() < 400; p++) {
        char c = text[p];
        if (c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == ',') break;
        if (c == 'd' || c == 'D' || c == 'q' || c == 'Q') c = 'E';
        tk.push_back(c);
      }
      // a bare-sign exponent ("1.5-03"): insert the E
      for (size_t p = 1; p < tk.size(); p++)
        if ((tk[p] == '+' || tk[p] == '-') && tk[p - 1] != 'E' && tk[p - 1] != 'e') { tk.insert(p, "E"); break; }
      double v = strtod(tk.c_str(), nullptr);
      if (divisor != 1.0) v = v / divisor;
      fval[q] = v;
    }
    DevBuf b_val;
    C2G_CUDA(ctx, b_val.alloc(ctx, sizeof(double) * nfb));
    C2G_CUDA(ctx, cudaMemcpyAsync(b_val.p, fval.data(), sizeof(double) * nfb, cudaMemcpyHostToDevice, st));
    k_patch<<<c2g_blocks_for(nfb, 256), 256, 0, st>>>(nfb, b_fbt.as<long long>(), b_val.as<double>(), ctx->grids[*handle].d);
    C2G_KERNEL_CHECK(ctx);
    C2G_CUDA(ctx, cudaStreamSynchronize(st));
  }
  if (consumed) *consumed = (size_t)hctl[2];
  if (nhost) *nhost = nfb;
  ctx->prof_collect();
  hg.ok = true;
  return C2G_OK;
}
