// seqio.hpp - aligned FASTA input and column masking of the host layer.
//   reading / upper-casing        src/seq/MSA_Stream.cpp:8-45 (genesis FastaReader, to_upper)
//   all-gap column mask           src/seq/MSA_Info.hpp:22-111 (gap characters "NOX.-?" for both data types)
//   or-mask of reference + query  src/main.cpp:470-494
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace epa_host {

struct Alignment {
  std::vector<std::string> names;
  std::vector<uint8_t> rows;     // [names.size()][sites], upper case
  size_t sites = 0;
  size_t size() const { return names.size(); }
  const uint8_t * row(size_t i) const { return rows.data() + i * sites; }
};

Alignment read_fasta(const std::string & path);                       // throws std::runtime_error
// Query files in the reference's binary 4-bit format (src/io/Binary_Fasta.hpp:33-96,252-310,
// src/io/encoding.hpp): decoded back to upper-case rows; DNA only, like the reference's converter.
bool is_bfast(const std::string & path);
Alignment read_bfast(const std::string & path);
// the converter behind -c/--bfast (Binary_Fasta::fasta_to_bfast :214-246): writes <out_dir>/<file name>.bfast, returns its path
std::string write_bfast(const Alignment & a, const std::string & fasta_path, std::string out_dir);
Alignment read_alignment(const std::string & path);                  // bfast if the magic matches, FASTA otherwise
std::vector<uint8_t> gap_mask(const Alignment & a);                  // 1 = every sequence has a gap character
Alignment apply_mask(const Alignment & a, const std::vector<uint8_t> & drop);

}  // namespace epa_host
