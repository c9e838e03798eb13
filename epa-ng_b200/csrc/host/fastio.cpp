#include "fastio.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <thread>

namespace epa_host {

// ---- memory map -------------------------------------------------------------------------------
MappedFile::MappedFile(const std::string & path)
{
  const int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) throw std::runtime_error("Cannot open file: " + path);
  struct stat st;
  if (::fstat(fd, &st) != 0) { ::close(fd); throw std::runtime_error("Cannot read file: " + path); }
  size_ = (size_t) st.st_size;
  if (size_)
  {
    void * p = ::mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { ::close(fd); throw std::runtime_error("Cannot map file: " + path); }
    ::madvise(p, size_, MADV_SEQUENTIAL);
    data_ = static_cast<const char *>(p);
  }
  ::close(fd);
}

MappedFile::~MappedFile()
{
  if (data_) ::munmap(const_cast<char *>(data_), size_);
}

// ---- helpers ----------------------------------------------------------------------------------
namespace {

struct Tables {
  bool space[256] = {};
  bool gap[256] = {};
  uint8_t upper[256];
  Tables()
  {
    for (const char * p = " \t\n\v\f\r"; *p; ++p) space[(unsigned char) *p] = true;
    for (const char * p = "NOX.-?nox"; *p; ++p) gap[(unsigned char) *p] = true;       // src/seq/MSA_Info.hpp:93-111
    for (int c = 0; c < 256; ++c) upper[c] = (uint8_t) ((c >= 'a' && c <= 'z') ? c - 32 : c);
  }
};
const Tables kT;

const char kBfastMagic[7] = {'B', 'F', 'A', 'S', 'T', '\0', '\0'};
const char kNtMap[17] = "-TGKCYSBAWRDMHVN";       // src/util/maps.hpp:9-14
struct NibblePairs {
  uint16_t t[256];                                  // the two characters of a packed byte, in memory order
  NibblePairs()
  {
    for (int b = 0; b < 256; ++b)
    {
      const unsigned char two[2] = {(unsigned char) kNtMap[b >> 4], (unsigned char) kNtMap[b & 15]};
      std::memcpy(&t[b], two, 2);
    }
  }
};
const NibblePairs kPairs;

// runs fn(t) on `threads` threads and rethrows the first exception
template <class F>
void parallel(int threads, F && fn)
{
  threads = std::max(1, threads);
  if (threads == 1) { fn(0); return; }
  std::vector<std::thread> pool;
  std::vector<std::exception_ptr> err((size_t) threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&, t]() { try { fn(t); } catch (...) { err[(size_t) t] = std::current_exception(); } });
  for (auto & th : pool) th.join();
  for (auto & e : err) if (e) std::rethrow_exception(e);
}

std::string record_name(const QueryRecord & r) { return std::string(r.name, r.name_len); }

// number of white-space bytes (the six of kT.space: 9..13 and 32) in [p, e): eight bytes per step with exact
// per-byte flags - no carries between bytes: for b < 0x80, (b & 0x7f) + K sets bit 7 iff b >= 0x80 - K
size_t count_space(const char * p, const char * e)
{
  size_t n = 0;
#if defined(__SSE2__)
  {
    const __m128i c9 = _mm_set1_epi8(9), c13 = _mm_set1_epi8(13), c32 = _mm_set1_epi8(32);
    for (; p + 16 <= e; p += 16)
    {
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p));
      const __m128i in = _mm_and_si128(_mm_cmpeq_epi8(_mm_max_epu8(c, c9), c), _mm_cmpeq_epi8(_mm_min_epu8(c, c13), c));
      n += (size_t) __builtin_popcount((unsigned) _mm_movemask_epi8(_mm_or_si128(in, _mm_cmpeq_epi8(c, c32))));
    }
  }
#endif
  const uint64_t L7 = 0x7f7f7f7f7f7f7f7full, H = 0x8080808080808080ull;
  for (; p + 8 <= e; p += 8)
  {
    uint64_t x;
    std::memcpy(&x, p, 8);
    const uint64_t lo = x & L7;
    const uint64_t ge9 = (lo + 0x7777777777777777ull);            // bit 7: low 7 bits >= 9
    const uint64_t ge14 = (lo + 0x7272727272727272ull);           // bit 7: low 7 bits >= 14
    const uint64_t y = lo ^ 0x2020202020202020ull;                // zero where the low 7 bits are 0x20
    const uint64_t eq32 = ~(y + L7);                              // bit 7 set where y == 0 (y <= 0x7f: no carry out)
    const uint64_t ws = ((ge9 & ~ge14) | eq32) & ~x & H;          // ... and the byte itself is below 0x80
    n += (size_t) __builtin_popcountll(ws);
  }
  for (; p < e; ++p) n += kT.space[(unsigned char) *p] ? 1 : 0;
  return n;
}

void index_fasta(const MappedFile & file, const std::string & path, int threads, bool want_mask, QueryIndex & idx)
{
  const char * d = file.data();
  const size_t n = file.size();
  // the first non-blank line must be a header (read_fasta: "sequence data before the first '>' line")
  {
    size_t p = 0;
    while (p < n && kT.space[(unsigned char) d[p]]) ++p;
    if (p == n) throw std::runtime_error(path + " contains no sequences");
    if (d[p] != '>' || (p > 0 && d[p - 1] != '\n')) throw std::runtime_error(path + ": sequence data before the first '>' line");
  }
  // 1. record starts: '>' at the beginning of a line, found by every thread in its byte range
  std::vector<std::vector<size_t>> starts((size_t) threads);
  parallel(threads, [&](int t)
  {
    const size_t lo = n * (size_t) t / (size_t) threads, hi = n * (size_t) (t + 1) / (size_t) threads;
    auto & v = starts[(size_t) t];
    const char * p = d + lo;
    const char * end = d + hi;
    while (p < end)
    {
      p = static_cast<const char *>(std::memchr(p, '>', (size_t) (end - p)));
      if (!p) break;
      if (p == d || p[-1] == '\n') v.push_back((size_t) (p - d));
      ++p;
    }
  });
  size_t total = 0;
  for (auto & v : starts) total += v.size();
  idx.records.resize(total);
  {
    size_t k = 0;
    for (auto & v : starts)
      for (size_t off : v)
      {
        QueryRecord & r = idx.records[k++];
        r.name = d + off + 1;
        r.seq = nullptr;
        r.seq_end = nullptr;
        r.name_len = 0;
        (void) r;
      }
    // header extents and sequence ranges
    parallel(threads, [&](int t)
    {
      const size_t lo = total * (size_t) t / (size_t) threads, hi = total * (size_t) (t + 1) / (size_t) threads;
      for (size_t i = lo; i < hi; ++i)
      {
        QueryRecord & r = idx.records[i];
        const char * stop = i + 1 < total ? idx.records[i + 1].name - 1 : d + n;
        const char * eol = static_cast<const char *>(std::memchr(r.name, '\n', (size_t) (stop - r.name)));
        if (!eol) eol = stop;
        const char * e = eol;
        while (e > r.name && (e[-1] == '\r' || e[-1] == ' ' || e[-1] == '\t')) --e;
        r.name_len = (uint32_t) (e - r.name);
        r.seq = eol < stop ? eol + 1 : stop;
        r.seq_end = stop;
      }
    });
  }
  if (total == 0) throw std::runtime_error(path + " contains no sequences");
  // 2. width of the first record, then every record is checked against it while the mask is built
  {
    size_t k = 0;
    for (const char * p = idx.records[0].seq; p < idx.records[0].seq_end; ++p)
      if (!kT.space[(unsigned char) *p]) ++k;
    idx.sites = k;
  }
  const size_t sites = idx.sites;
  std::vector<std::vector<uint8_t>> masks((size_t) threads);
  parallel(threads, [&](int t)
  {
    const size_t lo = total * (size_t) t / (size_t) threads, hi = total * (size_t) (t + 1) / (size_t) threads;
    std::vector<uint8_t> & m = masks[(size_t) t];
    m.assign(sites, 1);
    size_t open = want_mask ? sites : 0;          // columns still all-gap in this thread's records
    for (size_t i = lo; i < hi; ++i)
    {
      const QueryRecord & r = idx.records[i];
      size_t k = 0;
      if (open)
      {
        for (const char * p = r.seq; p < r.seq_end; ++p)
        {
          const unsigned char c = (unsigned char) *p;
          if (kT.space[c]) continue;
          if (k < sites && m[k] && !kT.gap[c]) { m[k] = 0; --open; }
          ++k;
        }
      }
      else
        k = (size_t) (r.seq_end - r.seq) - count_space(r.seq, r.seq_end);
      if (k != sites)
        throw std::runtime_error(path + " does not contain equal size sequences! First offending sequence: " + record_name(r));
    }
    if (!want_mask) std::fill(m.begin(), m.end(), (uint8_t) 0);
  });
  idx.gap_mask.assign(sites, want_mask ? 1 : 0);
  if (want_mask)
    for (auto & m : masks)
      if (!m.empty())
        for (size_t s = 0; s < sites; ++s) idx.gap_mask[s] &= m[s];
}

void index_bfast(const MappedFile & file, const std::string & path, int threads, QueryIndex & idx)
{
  const char * d = file.data();
  const size_t n = file.size();
  size_t pos = sizeof kBfastMagic;
  auto need = [&](size_t k) { if (pos + k > n) throw std::runtime_error(path + ": truncated bfast file"); };
  auto u64 = [&]() { need(8); uint64_t v; std::memcpy(&v, d + pos, 8); pos += 8; return v; };
  const uint64_t n_seq = u64();
  const uint64_t mask_len = u64();
  need(mask_len);
  idx.gap_mask.resize(mask_len);
  for (uint64_t i = 0; i < mask_len; ++i) idx.gap_mask[i] = d[pos + i] == '1';     // Binary_Fasta.hpp:60-66
  pos += mask_len;
  if (n_seq > (n - pos) / 16) throw std::runtime_error(path + ": truncated bfast file");
  const char * table = d + pos;                    // random-access table: (sequence id, byte offset) per entry, :53-64
  pos += n_seq * 16;
  idx.records.resize(n_seq);
  if (n_seq == 0) throw std::runtime_error(path + ": no sequences");
  // Entries lie back to back in table order; every thread parses a range of them from the table's offsets (the
  // pages of the map are then touched by all threads), and every entry must end where the next one starts.
  auto parse = [&](uint64_t i, size_t at, size_t & end)
  {
    auto rd = [&](size_t where) { if (where + 8 > n) throw std::runtime_error(path + ": truncated bfast file"); uint64_t v; std::memcpy(&v, d + where, 8); return v; };
    QueryRecord & r = idx.records[i];
    const uint64_t label_len = rd(at);
    if (label_len > n - (at + 8)) throw std::runtime_error(path + ": truncated bfast file");
    r.name = d + at + 8; r.name_len = (uint32_t) label_len;
    const size_t cat = at + 8 + (size_t) label_len;
    const uint64_t n_chars = rd(cat);
    const size_t packed = (size_t) ((n_chars + 1) / 2);
    if (packed > n - (cat + 8)) throw std::runtime_error(path + ": truncated bfast file");
    r.seq = d + cat + 8; r.seq_end = r.seq + packed;
    // (the width travels in seq_end - seq; odd and even widths are told apart by the first record's count)
    end = cat + 8 + packed;
    return n_chars;
  };
  size_t end0 = 0;
  {
    uint64_t off0; std::memcpy(&off0, table + 8, 8);
    if (off0 != pos) throw std::runtime_error(path + ": corrupt bfast offset table");
    idx.sites = parse(0, (size_t) off0, end0);
  }
  const uint64_t sites = idx.sites;
  parallel(threads, [&](int t)
  {
    const uint64_t lo = std::max<uint64_t>(1, n_seq * (uint64_t) t / (uint64_t) threads), hi = n_seq * (uint64_t) (t + 1) / (uint64_t) threads;
    for (uint64_t i = lo; i < hi; ++i)
    {
      uint64_t off; std::memcpy(&off, table + i * 16 + 8, 8);
      if (off >= n) throw std::runtime_error(path + ": truncated bfast file");
      size_t end;
      const uint64_t n_chars = parse(i, (size_t) off, end);
      if (n_chars != sites)
        throw std::runtime_error(path + " does not contain equal size sequences! First offending sequence: " + record_name(idx.records[i]));
    }
  });
  // back to back: every entry starts where the previous one ends
  for (uint64_t i = 1; i < n_seq; ++i)
    if (idx.records[i].name - 8 != idx.records[i - 1].seq_end) throw std::runtime_error(path + ": corrupt bfast offset table");
  (void) end0;
  if (idx.records.empty()) throw std::runtime_error(path + ": no sequences");
  if (idx.gap_mask.size() != idx.sites) idx.gap_mask.assign(idx.sites, 0);
}

}  // namespace

QueryIndex index_queries(const MappedFile & file, const std::string & path, int threads, bool want_mask)
{
  QueryIndex idx;
  idx.bfast = file.size() >= sizeof kBfastMagic && std::memcmp(file.data(), kBfastMagic, sizeof kBfastMagic) == 0;
  if (idx.bfast)
  {
    index_bfast(file, path, threads, idx);
    if (!want_mask) std::fill(idx.gap_mask.begin(), idx.gap_mask.end(), (uint8_t) 0);
  }
  else
    index_fasta(file, path, std::max(1, threads), want_mask, idx);
  return idx;
}

void decode_rows(const QueryIndex & idx, size_t first, size_t count, const std::vector<uint32_t> & keep, uint8_t * out, int threads)
{
  const size_t sites = idx.sites, width = keep.size();
  const bool all = width == sites;
  threads = (int) std::max<size_t>(1, std::min<size_t>((size_t) std::max(1, threads), count / 256 + 1));
  parallel(threads, [&](int t)
  {
    const size_t lo = count * (size_t) t / (size_t) threads, hi = count * (size_t) (t + 1) / (size_t) threads;
    std::vector<uint8_t> tmp(all ? 0 : sites);
    for (size_t i = lo; i < hi; ++i)
    {
      const QueryRecord & r = idx.records[first + i];
      uint8_t * row = out + i * width;
      uint8_t * dst = all ? row : tmp.data();
      if (idx.bfast)
      {
        const unsigned char * p = reinterpret_cast<const unsigned char *>(r.seq);
        size_t k = 0;
        for (; k + 1 < sites; k += 2, ++p) std::memcpy(dst + k, &kPairs.t[*p], 2);      // two characters per packed byte
        if (k < sites) dst[k] = (uint8_t) kNtMap[*p >> 4];
      }
      else
      {
        // one unwrapped line is the common case: a straight upper-casing copy (a loop the compiler vectorises) that
        // also notices white space; wrapped or padded records take the byte-wise path
        const char * p = r.seq;
        size_t k = 0;
        if ((size_t) (r.seq_end - r.seq) >= sites)
        {
          const unsigned char * src = reinterpret_cast<const unsigned char *>(r.seq);
          unsigned char odd = 0;
          size_t i = 0;
#if defined(__SSE2__)
          {
            const __m128i sp = _mm_set1_epi8(' '), ca = _mm_set1_epi8('a'), cz = _mm_set1_epi8('z'), d32 = _mm_set1_epi8(32);
            __m128i oddv = _mm_setzero_si128();
            for (; i + 16 <= sites; i += 16)
            {
              const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
              oddv = _mm_or_si128(oddv, _mm_cmpeq_epi8(_mm_min_epu8(c, sp), c));                     // c <= ' '
              const __m128i low = _mm_and_si128(_mm_cmpeq_epi8(_mm_max_epu8(c, ca), c), _mm_cmpeq_epi8(_mm_min_epu8(c, cz), c));
              _mm_storeu_si128(reinterpret_cast<__m128i *>(dst + i), _mm_sub_epi8(c, _mm_and_si128(low, d32)));
            }
            odd = _mm_movemask_epi8(oddv) ? 1 : 0;
          }
#endif
          for (; i < sites; ++i)
          {
            const unsigned char c = src[i];
            odd |= (unsigned char) (c <= ' ');
            dst[i] = (uint8_t) (c - (((unsigned char) (c - 'a') < 26) ? 32 : 0));
          }
          if (!odd) k = sites;
        }
        while (p < r.seq_end && k < sites)
        {
          const unsigned char c = (unsigned char) *p++;
          if (kT.space[c]) continue;
          dst[k++] = kT.upper[c];
        }
      }
      if (!all)
        for (size_t j = 0; j < width; ++j) row[j] = tmp[keep[j]];
    }
  });
}

// ---- printf-exact fixed-point formatting ------------------------------------------------------
namespace {
const char kDigits2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
    "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
    "8081828384858687888990919293949596979899";

const uint64_t kPow10[19] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                             1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                             100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
                             1000000000000000000ull};

// decimal digits of v, most significant first; returns the count
inline size_t put_u64(char * out, uint64_t v)
{
  char buf[20];
  size_t n = 0;
  while (v >= 100) { const unsigned r = (unsigned) (v % 100); v /= 100; buf[n++] = kDigits2[2 * r + 1]; buf[n++] = kDigits2[2 * r]; }
  if (v >= 10) { buf[n++] = kDigits2[2 * v + 1]; buf[n++] = kDigits2[2 * v]; }
  else buf[n++] = (char) ('0' + v);
  for (size_t i = 0; i < n; ++i) out[i] = buf[n - 1 - i];
  return n;
}

// exactly `width` digits of v (v < 10^width), zero padded
inline void put_u64_padded(char * out, uint64_t v, int width)
{
  int i = width;
  while (i >= 2) { const unsigned r = (unsigned) (v % 100); v /= 100; i -= 2; out[i] = kDigits2[2 * r]; out[i + 1] = kDigits2[2 * r + 1]; }
  if (i == 1) out[0] = (char) ('0' + v % 10);
}
}  // namespace

size_t format_fixed(char * out, double x, int precision)
{
  uint64_t bits;
  std::memcpy(&bits, &x, 8);
  const bool neg = (bits >> 63) != 0;
  const double ax = std::fabs(x);
  if (!(ax < 1e15) || precision < 0 || precision > 18)
    return (size_t) std::snprintf(out, 48 + (size_t) std::max(0, precision) + 310, "%.*f", precision, x);
  const int be = (int) ((bits >> 52) & 0x7ff);
  uint64_t m = bits & 0xfffffffffffffull;
  int e;
  if (be == 0) e = -1074;
  else { m |= 1ull << 52; e = be - 1075; }
  // ax = m * 2^e with e < 0 (ax < 2^50); q = round_half_even(m * 10^precision / 2^-e)
  const unsigned __int128 P = (unsigned __int128) m * kPow10[precision];
  const int k = -e;
  unsigned __int128 q;
  if (k > 127) q = 0;
  else
  {
    q = P >> k;
    const unsigned __int128 rem = P & ((((unsigned __int128) 1) << k) - 1);
    const unsigned __int128 half = ((unsigned __int128) 1) << (k - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
  }
  char * p = out;
  if (neg) *p++ = '-';
  const uint64_t scale = kPow10[precision];
  uint64_t ip, fp;
  if ((q >> 64) == 0) { const uint64_t q64 = (uint64_t) q; ip = q64 / scale; fp = q64 % scale; }
  else { ip = (uint64_t) (q / scale); fp = (uint64_t) (q % scale); }
  p += put_u64(p, ip);
  if (precision > 0)
  {
    *p++ = '.';
    put_u64_padded(p, fp, precision);
    p += precision;
  }
  return (size_t) (p - out);
}

void json_escape(std::string & out, const char * s, size_t n)
{
  // the common case has nothing to escape: one scan, one append
  size_t i = 0;
  for (; i < n; ++i)
  {
    const unsigned char c = (unsigned char) s[i];
    if (c == '"' || c == '\\' || c < 0x20) break;
  }
  out.append(s, i);
  for (; i < n; ++i)
  {
    const unsigned char c = (unsigned char) s[i];
    if (c == '"' || c == '\\') { out += '\\'; out += (char) c; }
    else if (c < 0x20)
    {
      char buf[8];
      std::snprintf(buf, sizeof buf, "\\u%04x", c);
      out += buf;
    }
    else out += (char) c;
  }
}

void append_pquery(std::string & out, const char * name, size_t name_len, const PlacementFields * p, uint32_t count,
                   int precision, bool last)
{
  char buf[2048];
  out += "    {\"p\": [\n";
  for (uint32_t k = 0; k < count; ++k)
  {
    char * w = buf;
    std::memcpy(w, "      [", 7); w += 7;
    w += put_u64(w, p[k].branch_id);
    *w++ = ','; *w++ = ' ';
    w += format_fixed(w, p[k].likelihood, precision); *w++ = ','; *w++ = ' ';
    w += format_fixed(w, p[k].lwr, precision); *w++ = ','; *w++ = ' ';
    w += format_fixed(w, p[k].distal_length, precision); *w++ = ','; *w++ = ' ';
    w += format_fixed(w, p[k].pendant_length, precision);
    *w++ = ']';
    if (k + 1 < count) *w++ = ',';
    *w++ = '\n';
    out.append(buf, (size_t) (w - buf));
  }
  out += "      ],\n    \"n\": [\"";
  json_escape(out, name, name_len);
  out += "\"]\n    }";
  out += last ? "\n" : ",\n";
}

}  // namespace epa_host
