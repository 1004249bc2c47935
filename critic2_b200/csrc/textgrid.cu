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
//     high word is the correctly rounded result unless the low word is within that error of a rounding boundary;
//   * those tokens, subnormal results, overflows and mantissas longer than 19 digits take an exact multi-word
//     integer comparison of the decimal value with the midpoints around a candidate (decimal_exact), also on the
//     device (never needed on cube / CHGCAR data; the tests force it).  Nothing is converted on the host.
// Separators: blank, tab, newline, carriage return, comma.  Exponent letters E, D, Q (either case) or a bare sign
// ("1.5-03").  The r*c repeat form of list-directed input is not supported (neither format writes it).
#include "common.cuh"
#include "group.h"

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
  int* nfb;                  // tokens that took the exact big-integer path
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

// ------------------------------------------------------------------------------------------------
// Exact conversion for the tokens the fast paths cannot decide (result within 2^-98 of a rounding boundary,
// subnormal or huge results, mantissas longer than 19 digits): the decimal value D = mant * 10^q, with ALL digits of
// the token in a multi-word integer, is compared exactly with the midpoints around a candidate double, and the
// candidate is moved until D lies between them; ties go to the even mantissa (IEEE round-to-nearest-even, what
// strtod and a list-directed READ do).  Rare (never on cube / CHGCAR data), so it is written for clarity, not speed.
// ------------------------------------------------------------------------------------------------
constexpr int BIGW = 56;  // 1792 bits: 60 digits (200 bits) * 5^400 (929 bits) and shifts of up to ~1100 bits never meet
struct Big {
  unsigned w[BIGW];
  int n;  // words in use (>= 1)
};
__device__ void big_set(Big& b, unsigned long long v) {
  b.w[0] = (unsigned)v; b.w[1] = (unsigned)(v >> 32);
  b.n = b.w[1] ? 2 : 1;
}
__device__ void big_mul_add(Big& b, unsigned k, unsigned add) {  // b = b * k + add
  unsigned long long carry = add;
  for (int i = 0; i < b.n; i++) {
    const unsigned long long t = (unsigned long long)b.w[i] * k + carry;
    b.w[i] = (unsigned)t;
    carry = t >> 32;
  }
  if (carry && b.n < BIGW) b.w[b.n++] = (unsigned)carry;
}
__device__ void big_mul_pow5(Big& b, int e) {
  for (; e >= 13; e -= 13) big_mul_add(b, 1220703125u, 0);  // 5^13
  for (; e > 0; e--) big_mul_add(b, 5u, 0);
}
__device__ void big_shl(Big& b, int s) {
  const int ws = s >> 5, bs = s & 31;
  if (b.n == 1 && b.w[0] == 0) return;
  int nn = b.n + ws + 1;
  if (nn > BIGW) nn = BIGW;
  for (int i = nn - 1; i >= 0; i--) {
    const int src = i - ws;
    unsigned lo = (src >= 0 && src < b.n) ? b.w[src] : 0u;
    unsigned below = (src - 1 >= 0 && src - 1 < b.n) ? b.w[src - 1] : 0u;
    b.w[i] = bs ? ((lo << bs) | (below >> (32 - bs))) : lo;
  }
  b.n = nn;
  while (b.n > 1 && b.w[b.n - 1] == 0) b.n--;
}
__device__ int big_cmp(const Big& a, const Big& b) {
  if (a.n != b.n) return a.n > b.n ? 1 : -1;
  for (int i = a.n - 1; i >= 0; i--)
    if (a.w[i] != b.w[i]) return a.w[i] > b.w[i] ? 1 : -1;
  return 0;
}
// sign of  mant * 10^q  -  t * 2^e2
__device__ int cmp_decimal_binary(const Big& mant, int q, unsigned long long t, int e2) {
  Big L = mant, R;
  big_set(R, t);
  if (q > 0) big_mul_pow5(L, q);
  if (q < 0) big_mul_pow5(R, -q);
  const int d = q - e2;  // net power of two on the left
  if (d >= 0) big_shl(L, d); else big_shl(R, -d);
  return big_cmp(L, R);
}
__device__ __forceinline__ void dbl_decompose(double c, unsigned long long& M, int& E) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(c);
  const int ex = (int)((bits >> 52) & 0x7ff);
  const unsigned long long fr = bits & ((1ull << 52) - 1);
  if (ex == 0) { M = fr; E = -1074; } else { M = fr | (1ull << 52); E = ex - 1075; }
}
// tok: the token (sign already consumed by the caller or not -- handled here), terminated by a separator
__device__ __noinline__ double decimal_exact(const unsigned char* tok) {
  int p = 0;
  if (tok[p] == '+' || tok[p] == '-') p++;
  Big mant;
  big_set(mant, 0);
  int nd = 0, frac = 0;
  bool dot = false;
  unsigned long long m19 = 0;
  int n19 = 0;
  for (;; p++) {
    const unsigned char ch = tok[p];
    if (ch >= '0' && ch <= '9') {
      if (nd > 0 || ch != '0') {
        big_mul_add(mant, 10u, (unsigned)(ch - '0'));
        nd++;
        if (n19 < 19) { m19 = m19 * 10ull + (unsigned)(ch - '0'); n19++; }
      }
      if (dot) frac++;
    } else if (ch == '.' && !dot) dot = true;
    else break;
  }
  int ex = 0;
  {
    unsigned char ch = tok[p];
    if (ch == 'e' || ch == 'E' || ch == 'd' || ch == 'D' || ch == 'q' || ch == 'Q' || ch == '+' || ch == '-') {
      if (!(ch == '+' || ch == '-')) ch = tok[++p];
      bool eneg = false;
      if (ch == '+' || ch == '-') { eneg = ch == '-'; ch = tok[++p]; }
      while (ch >= '0' && ch <= '9') { if (ex < 100000) ex = ex * 10 + (ch - '0'); ch = tok[++p]; }
      if (eneg) ex = -ex;
    }
  }
  if (nd == 0) return 0.0;
  const int q = ex - frac;
  const int dexp = q + nd - 1;  // decimal exponent of the leading digit
  if (dexp > 309) return __longlong_as_double(0x7ff0000000000000ll);
  if (dexp < -326) return 0.0;
  // candidate from plain doubles (a few ulp off at worst; the loop below repairs it)
  double c = (double)m19;
  int q19 = q + (nd - n19);
  while (q19 > 0) { const int s = q19 > 300 ? 300 : q19; c *= c2g_pow10_hi[s]; q19 -= s; }
  while (q19 < 0) { const int s = -q19 > 300 ? 300 : -q19; c /= c2g_pow10_hi[s]; q19 += s; }
  const double dmax = __longlong_as_double(0x7fefffffffffffffll);
  if (!(c <= dmax)) c = dmax;
  for (int it = 0; it < 128; it++) {
    unsigned long long M;
    int E;
    dbl_decompose(c, M, E);
    // D against the midpoint above c: (2M+1) * 2^(E-1)
    const int up = cmp_decimal_binary(mant, q, 2 * M + 1, E - 1);
    if (up > 0 || (up == 0 && (M & 1))) {
      if (c == dmax) return __longlong_as_double(0x7ff0000000000000ll);
      c = __longlong_as_double(__double_as_longlong(c) + 1);
      if (up == 0) return c;  // tie resolved to the even neighbour
      continue;
    }
    if (up == 0) return c;
    if (c == 0.0) return c;
    const double pr = __longlong_as_double(__double_as_longlong(c) - 1);
    unsigned long long Mp;
    int Ep;
    dbl_decompose(pr, Mp, Ep);
    const int dn = cmp_decimal_binary(mant, q, 2 * Mp + 1, Ep - 1);
    if (dn < 0) { c = pr; continue; }
    if (dn == 0) return (Mp & 1) ? c : pr;
    return c;
  }
  return c;
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
    if (!ok) {  // exact big-integer path (counted)
      v = decimal_exact(s_b + pstart);
      atomicAdd(A.nfb, 1);
    }
    if (neg) v = -v;
    if (A.divisor != 1.0) v = v / A.divisor;
    A.out[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Formatted output: the value loops of writegrid_cube (crystalmod@write.f90:3556-3565, formats (1p,6(" ",E12.5E3)) and
// (6(" ",E22.14E3))) and of NCIPLOT's write_cube_body (nci@proc.f90:916-932, (6(" ",1p,e13.5e3))).  One thread per
// value: S significant digits, CORRECTLY ROUNDED (nearest, ties to even -- what the Fortran run-time library prints),
// laid out by the rules of the Ew.dEe edit descriptor with scale factor k (Fortran 2018 13.7.2.3.3): k = 1:
// [-]d.ddd..E+eee with d digits after the point, k = 0: [-]0.ddd..E+eee; right-justified in w characters; a field
// that does not fit (e.g. a NEGATIVE value in 1p,E12.5E3, which needs 13 characters) is w asterisks, like the
// reference prints it.  Digits: |x| * 10^p in double-double; if the fraction is within 2^-96 of one half the decision
// is made exactly with the multi-word comparison above.
// ------------------------------------------------------------------------------------------------
struct FormatArgs {
  const double* f;
  int n1, n2, n3;        // resident grid shape
  int layout;            // 0: rows along index 1 as stored; 1: cube order (rows along index 3) with ishift
  int s1, s2, s3;
  int w, d, k;           // Ew.dE3 with scale factor k (0 or 1)
  long long nrows; int L;
  long long rowbytes;
  unsigned char* out;
};

__device__ __forceinline__ void dd_mul_d(double ah, double al, double bh, double bl, double& rh, double& rl) {
  double p, e;
  two_prod(ah, bh, p, e);
  e += ah * bl + al * bh;
  fast_two_sum(p, e, rh, rl);
}
// ax > 0 finite: N = the S leading digits, correctly rounded; e10 = decimal exponent of the first digit
__device__ void dec_digits(double ax, int S, unsigned long long& N, int& e10) {
  unsigned long long M;
  int E;
  dbl_decompose(ax, M, E);
  int e = (int)floor(log10(ax));
  unsigned long long p10S = 1;
  for (int q = 0; q < S; q++) p10S *= 10ull;
  for (int it = 0; it < 6; it++) {
    const int p = S - 1 - e;
    double yh, yl;
    if (p >= 0) {
      const int pa = p > C2G_POW10_MAX ? C2G_POW10_MAX : p;
      dd_mul_d(ax, 0.0, c2g_pow10_hi[pa], c2g_pow10_lo[pa], yh, yl);
      if (p > pa) dd_mul_d(yh, yl, c2g_pow10_hi[p - pa], c2g_pow10_lo[p - pa], yh, yl);
    } else {
      const double ph = c2g_pow10_hi[-p], pl = c2g_pow10_lo[-p];
      const double q1 = ax / ph;
      double pp, ee;
      two_prod(q1, ph, pp, ee);
      const double r = ((ax - pp) - ee) - q1 * pl;
      fast_two_sum(q1, r / ph, yh, yl);
    }
    double nf = floor(yh);
    double frac = (yh - nf) + yl;
    if (frac < 0.0) { nf -= 1.0; frac += 1.0; }
    if (frac >= 1.0) { nf += 1.0; frac -= 1.0; }
    const unsigned long long n = (unsigned long long)nf;
    if (n >= p10S) { e++; continue; }            // the estimate of the exponent was one too small
    if (n < p10S / 10ull) { e--; continue; }     // ... or one too large
    unsigned long long Nr;
    if (fabs(frac - 0.5) <= 1.3e-29 * (nf + 1.0)) {  // too close to call: exact comparison of n + 1/2 with ax * 10^p
      Big mant;
      big_set(mant, 2ull * n + 1ull);
      const int c = cmp_decimal_binary(mant, -p, M, E + 1);
      Nr = c > 0 ? n : (c < 0 ? n + 1 : n + (n & 1ull));
    } else {
      Nr = frac > 0.5 ? n + 1 : n;
    }
    if (Nr >= p10S) { Nr = p10S / 10ull; e++; }  // 9.99..96 rounds up to 10.0..0
    N = Nr; e10 = e;
    return;
  }
  N = p10S / 10ull; e10 = e;  // not reached
}

__global__ void __launch_bounds__(256) k_format(const __grid_constant__ FormatArgs A) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.nrows * A.L) return;
  const long long r = t / A.L;
  const int c = (int)(t % A.L);
  size_t src;
  if (A.layout == 0) src = (size_t)t;
  else {
    const int iix = (int)(r / A.n2), iiy = (int)(r % A.n2);
    const int ix = (iix + A.s1) % A.n1, iy = (iiy + A.s2) % A.n2, iz = (c + A.s3) % A.n3;
    src = (size_t)ix + (size_t)A.n1 * ((size_t)iy + (size_t)A.n2 * iz);
  }
  const double x = A.f[src];
  const int w = A.w, S = A.d + (A.k == 1 ? 1 : 0);
  char buf[32];
  int len = 0;
  const long long bits = __double_as_longlong(x);
  const bool neg = bits < 0;
  const int ex = (int)((bits >> 52) & 0x7ff);
  if (ex == 0x7ff) {
    const bool isnan_ = (bits & ((1ll << 52) - 1)) != 0;
    if (isnan_) { buf[0] = 'N'; buf[1] = 'a'; buf[2] = 'N'; len = 3; }
    else {
      if (neg) buf[len++] = '-';
      const char* word = (w >= 8 + (neg ? 1 : 0)) ? "Infinity" : "Inf";
      for (int q = 0; word[q]; q++) buf[len++] = word[q];
    }
  } else {
    unsigned long long N = 0;
    int e10 = 0;
    if (x != 0.0) dec_digits(fabs(x), S, N, e10);
    int pexp = (x == 0.0) ? 0 : (A.k == 1 ? e10 : e10 + 1);
    char dig[20];
    for (int q = S - 1; q >= 0; q--) { dig[q] = (char)('0' + (int)(N % 10ull)); N /= 10ull; }
    const int body = (A.k == 1 ? 1 + 1 + (S - 1) : 2 + S) + 5;  // digits, point, E+eee
    bool lead0 = A.k == 0;
    int need = body + (neg ? 1 : 0);
    if (need > w && lead0) { lead0 = false; need--; }  // the zero before the point is optional
    if (need <= w) {
      if (neg) buf[len++] = '-';
      if (A.k == 1) { buf[len++] = dig[0]; buf[len++] = '.'; for (int q = 1; q < S; q++) buf[len++] = dig[q]; }
      else { if (lead0) buf[len++] = '0'; buf[len++] = '.'; for (int q = 0; q < S; q++) buf[len++] = dig[q]; }
      buf[len++] = 'E';
      buf[len++] = pexp < 0 ? '-' : '+';
      const int ae = pexp < 0 ? -pexp : pexp;
      buf[len++] = (char)('0' + ae / 100); buf[len++] = (char)('0' + (ae / 10) % 10); buf[len++] = (char)('0' + ae % 10);
    } else {
      len = w + 1;  // asterisks
    }
  }
  unsigned char* o = A.out + (size_t)r * A.rowbytes + (size_t)c * (w + 1) + c / 6;
  *o++ = ' ';
  if (len > w) { for (int q = 0; q < w; q++) o[q] = '*'; }
  else {
    const int pad = w - len;
    for (int q = 0; q < pad; q++) o[q] = ' ';
    for (int q = 0; q < len; q++) o[pad + q] = (unsigned char)buf[q];
  }
  // end of record: after the 6th value of a line the format reverts; when the data list ends inside the group
  // 6(" ",E...) the literal " " that precedes the next data edit descriptor is still written (Fortran format
  // control) -- a partial line therefore ends with a blank, as in the reference's own cube files
  // (tests/005_plot/ref/029_cube_precise_0[12].cube)
  if (c % 6 == 5) o[w] = '\n';
  else if (c == A.L - 1) { o[w] = ' '; o[w + 1] = '\n'; }
}

}  // namespace

// Parses the first n1*n2*n3 numbers of `text` (host memory, nbytes bytes) into a new resident grid.
extern "C" int c2g_grid_parse_text(c2g_context* ctx, const char* text, size_t nbytes, const int n[3], int order, double divisor,
                                   int* handle, size_t* consumed, long long* nslow) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_NOT_ON_GROUP(ctx, "c2g_grid_parse_text");
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
  DevBuf b_text, b_cnt, b_off, b_ctl;
  C2G_CUDA(ctx, b_text.alloc(ctx, padded));
  C2G_CUDA(ctx, cudaMemsetAsync((char*)b_text.p + nbytes, ' ', padded - nbytes, st));
  ctx->prof_begin("text_h2d");
  C2G_CUDA(ctx, cudaMemcpyAsync(b_text.p, text, nbytes, cudaMemcpyHostToDevice, st));
  ctx->prof_end(0);
  C2G_CUDA(ctx, b_cnt.alloc(ctx, sizeof(int) * nchunk));
  C2G_CUDA(ctx, b_off.alloc(ctx, sizeof(long long) * nchunk));
  C2G_CUDA(ctx, b_ctl.alloc(ctx, 64));
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
  if (consumed) *consumed = (size_t)hctl[2];
  if (nslow) *nslow = nfb;
  ctx->prof_collect();
  hg.ok = true;
  return C2G_OK;
}

// Formats a resident grid as the value block of a cube file.  out == NULL: only *nbytes (the size needed) is set.
extern "C" int c2g_grid_format_text(c2g_context* ctx, int handle, int layout, const int ishift[3], int width, int digits, int scale,
                                    char* out, size_t cap, size_t* nbytes) {
  if (!ctx) return C2G_ERR_ARG;
  C2G_NOT_ON_GROUP(ctx, "c2g_grid_format_text");
  if (!nbytes) return ctx->fail(C2G_ERR_ARG, "c2g_grid_format_text: null argument");
  if (handle < 0 || handle >= (int)ctx->grids.size() || !ctx->grids[handle].used)
    return ctx->fail(C2G_ERR_ARG, "c2g_grid_format_text: invalid grid handle %d", handle);
  if (layout != C2G_TEXT_ROWS_INDEX1 && layout != C2G_TEXT_ROWS_INDEX3) return ctx->fail(C2G_ERR_ARG, "c2g_grid_format_text: bad layout %d", layout);
  if ((scale != 0 && scale != 1) || digits < 1 || digits + scale > 17 || width < 1 || width > 30)
    return ctx->fail(C2G_ERR_ARG, "c2g_grid_format_text: unsupported edit descriptor E%d.%dE3 with %dP", width, digits, scale);
  c2g_grid_ready(ctx, handle);
  const c2g_grid& g = ctx->grids[handle];
  FormatArgs A;
  A.f = g.d; A.n1 = g.n[0]; A.n2 = g.n[1]; A.n3 = g.n[2]; A.layout = layout;
  A.s1 = A.s2 = A.s3 = 0;
  if (ishift && layout == C2G_TEXT_ROWS_INDEX3) {
    A.s1 = ((ishift[0] % A.n1) + A.n1) % A.n1; A.s2 = ((ishift[1] % A.n2) + A.n2) % A.n2; A.s3 = ((ishift[2] % A.n3) + A.n3) % A.n3;
  }
  A.w = width; A.d = digits; A.k = scale;
  if (layout == C2G_TEXT_ROWS_INDEX1) { A.L = A.n1; A.nrows = (long long)A.n2 * A.n3; }
  else { A.L = A.n3; A.nrows = (long long)A.n1 * A.n2; }
  A.rowbytes = (long long)A.L * (width + 1) + (A.L + 5) / 6 + (A.L % 6 != 0 ? 1 : 0);  // + the blank that ends a partial line
  const size_t total = (size_t)A.nrows * (size_t)A.rowbytes;
  *nbytes = total;
  if (!out || cap == 0) return C2G_OK;  // step 1: the size only
  if (cap < total) return ctx->fail(C2G_ERR_ARG, "c2g_grid_format_text: output buffer too small (%zu < %zu)", cap, total);
  DevBuf b_out;
  C2G_CUDA(ctx, b_out.alloc(ctx, total));
  A.out = (unsigned char*)b_out.p;
  const long long nval = A.nrows * A.L;
  ctx->prof_begin("text_format");
  k_format<<<(unsigned)((nval + 255) / 256), 256, 0, ctx->stream>>>(A);
  ctx->prof_end();
  C2G_KERNEL_CHECK(ctx);
  ctx->prof_begin("text_d2h");
  C2G_CUDA(ctx, cudaMemcpyAsync(out, b_out.p, total, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->prof_end(0);
  C2G_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->prof_collect();
  return C2G_OK;
}
