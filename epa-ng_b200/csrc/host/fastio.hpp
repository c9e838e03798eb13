// fastio.hpp - streaming, multi-threaded query input and jplace formatting for the files -> jplace
// pipeline (epa_run_files*). What the reference does with two helper threads
//   chunk prefetch          src/seq/MSA_Stream.cpp:79-85 (std::async read of the next chunk)
//   first pass over queries src/seq/MSA_Info.hpp:22-111 (sequence count, width, all-gap column mask)
//   jplace formatting       src/io/jplace_util.cpp:20-64, async write src/io/jplace_writer.hpp:58-69
// is done here with a memory-mapped query file, an index of its records built by all host threads,
// chunk decoding straight into pinned staging buffers, and a printf-exact fixed-point formatter.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace epa_host {

// read-only memory map of a file
class MappedFile {
 public:
  explicit MappedFile(const std::string & path);
  ~MappedFile();
  MappedFile(const MappedFile &) = delete;
  MappedFile & operator=(const MappedFile &) = delete;
  const char * data() const { return data_; }
  size_t size() const { return size_; }
 private:
  const char * data_ = nullptr;
  size_t size_ = 0;
};

// One record of a query file: label and where its characters start. FASTA: `seq` points at the first
// sequence line, `seq_end` at the next record (or the end of the file). bfast: `seq` points at the
// packed 4-bit codes.
struct QueryRecord {
  const char * name;
  uint32_t name_len;
  const char * seq;
  const char * seq_end;
};

// Index of an aligned query file (FASTA, or the reference's bfast): records in file order, alignment
// width, all-gap column mask over the "NOX.-?" characters. Built by `threads` threads over byte
// ranges of the map. Throws std::runtime_error with the reference's messages (unequal lengths, ...).
struct QueryIndex {
  bool bfast = false;
  size_t sites = 0;
  std::vector<QueryRecord> records;
  std::vector<uint8_t> gap_mask;            // [sites], 1 = every query has a gap character there
};
QueryIndex index_queries(const MappedFile & file, const std::string & path, int threads, bool want_mask);

// Decodes records [first, first + count) into upper-case rows out[count][keep.size()], keeping only
// the listed columns (the complement of the pre-mask); `threads` threads, rows split evenly.
void decode_rows(const QueryIndex & idx, size_t first, size_t count, const std::vector<uint32_t> & keep, uint8_t * out,
                 int threads);

// printf("%.*f") of a finite double, digit for digit (round-half-even on the exact binary value, as
// glibc does): 128-bit integer arithmetic for |x| < 1e15 and precision <= 18, snprintf otherwise.
// Returns the number of characters written (no terminator); `out` must hold 48 + precision bytes.
size_t format_fixed(char * out, double x, int precision);

// JSON string contents: '"', '\\' and control characters escaped
void json_escape(std::string & out, const char * s, size_t n);

// One placement record as the caller's 40-byte struct lays it out (epa_placement, include/epa_b200.h)
struct PlacementFields { uint64_t branch_id; double likelihood, lwr, pendant_length, distal_length; };

// appends the jplace text of one pquery (src/io/jplace_util.cpp:20-64: field order edge, logl, lwr, DISTAL, PENDANT)
void append_pquery(std::string & out, const char * name, size_t name_len, const PlacementFields * p, uint32_t count,
                   int precision, bool last);

}  // namespace epa_host
