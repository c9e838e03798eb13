#include "seqio.hpp"

#include <cctype>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace epa_host {

// ---- bfast ------------------------------------------------------------------------------------
// Layout (little endian, 64-bit integers), src/io/Binary_Fasta.hpp:33-96:
//   "BFAST\0\0" (7 bytes) | n_sequences | mask: length + '0'/'1' characters (all-gap columns) |
//   n x (sequence id, byte offset) | per sequence: label length + label, character count,
//   ceil(count / 2) bytes of two 4-bit codes each (first character in the high nibble; code =
//   index into "-TGKCYSBAWRDMHVN", src/util/maps.hpp:9-14; odd lengths are padded with code 0).
static const char kBfastMagic[7] = {'B', 'F', 'A', 'S', 'T', '\0', '\0'};
static const char kNtMap[17] = "-TGKCYSBAWRDMHVN";

static std::string slurp(const std::string & path)
{
  FILE * fh = std::fopen(path.c_str(), "rb");
  if (!fh) throw std::runtime_error("Cannot open file: " + path);
  std::fseek(fh, 0, SEEK_END);
  const long size = std::ftell(fh);
  std::fseek(fh, 0, SEEK_SET);
  std::string data((size_t) (size > 0 ? size : 0), '\0');
  const bool ok = size <= 0 || std::fread(&data[0], 1, (size_t) size, fh) == (size_t) size;
  std::fclose(fh);
  if (!ok) throw std::runtime_error("Cannot read file: " + path);
  return data;
}

bool is_bfast(const std::string & path)
{
  FILE * fh = std::fopen(path.c_str(), "rb");
  if (!fh) return false;
  char head[7] = {};
  const size_t got = std::fread(head, 1, sizeof head, fh);
  std::fclose(fh);
  return got == sizeof head && std::memcmp(head, kBfastMagic, sizeof head) == 0;
}

Alignment read_bfast(const std::string & path)
{
  const std::string data = slurp(path);
  size_t pos = 0;
  auto need = [&](size_t n) { if (pos + n > data.size()) throw std::runtime_error(path + ": truncated bfast file"); };
  auto u64 = [&]() { need(8); uint64_t v; std::memcpy(&v, data.data() + pos, 8); pos += 8; return v; };
  need(sizeof kBfastMagic);
  if (std::memcmp(data.data(), kBfastMagic, sizeof kBfastMagic) != 0) throw std::runtime_error("File is not an epa::Binary_Fasta file");
  pos = sizeof kBfastMagic;
  const uint64_t n_seq = u64();
  const uint64_t mask_len = u64();
  need(mask_len); pos += mask_len;                 // the stored all-gap mask is recomputed from the rows
  need(n_seq * 16); pos += n_seq * 16;             // random-access table: entries are read in file order
  Alignment a;
  a.names.reserve(n_seq);
  for (uint64_t i = 0; i < n_seq; ++i)
  {
    const uint64_t label_len = u64();
    need(label_len);
    a.names.emplace_back(data.substr(pos, label_len));
    pos += label_len;
    const uint64_t n_chars = u64();
    if (i == 0) { a.sites = n_chars; a.rows.reserve(n_seq * n_chars); }
    else if (n_chars != a.sites)
      throw std::runtime_error(path + " does not contain equal size sequences! First offending sequence: " + a.names.back());
    const size_t packed = (size_t) ((n_chars + 1) / 2);
    need(packed);
    for (uint64_t k = 0; k < n_chars; ++k)
    {
      const unsigned char byte = (unsigned char) data[pos + (k >> 1)];
      a.rows.push_back((uint8_t) kNtMap[(k & 1) ? (byte & 15) : (byte >> 4)]);
    }
    pos += packed;
  }
  if (a.names.empty()) throw std::runtime_error(path + ": no sequences");
  return a;
}

// Writer of the same format (Binary_Fasta::fasta_to_bfast, src/io/Binary_Fasta.hpp:33-70,214-246):
// header with the all-gap column mask and the random-access table, then label + packed sites per
// sequence. DNA only: a character outside "-TGKCYSBAWRDMHVN" is refused as the reference refuses
// amino-acid data (ensure_dna, :142-153).
std::string write_bfast(const Alignment & a, const std::string & fasta_path, std::string out_dir)
{
  int code[256];
  for (int & c : code) c = -1;
  for (int i = 0; i < 16; ++i) code[(unsigned char) kNtMap[i]] = i;
  for (uint8_t ch : a.rows)
    if (code[ch] < 0)
      throw std::runtime_error(std::string("AA DATA NOT SUPPORTED for conversion to bfast! Sorry! Offending char: ") + (char) ch);
  const size_t slash = fasta_path.find_last_of('/');
  if (!out_dir.empty() && out_dir.back() != '/') out_dir += '/';
  const std::string out_path = out_dir + (slash == std::string::npos ? fasta_path : fasta_path.substr(slash + 1)) + ".bfast";

  std::string out;
  auto put_u64 = [&](uint64_t v) { char b[8]; std::memcpy(b, &v, 8); out.append(b, 8); };
  out.append(kBfastMagic, sizeof kBfastMagic);
  put_u64(a.size());
  const std::vector<uint8_t> mask = gap_mask(a);
  put_u64(mask.size());
  for (uint8_t m : mask) out.push_back(m ? '1' : '0');
  const size_t packed = (a.sites + 1) / 2;
  uint64_t offset = sizeof kBfastMagic + 8 + a.size() * 16 + mask.size() + 8;      // data_section_offset, :33-36
  for (size_t i = 0; i < a.size(); ++i)
  {
    put_u64(i);
    put_u64(offset);
    offset += 16 + a.names[i].size() + packed;
  }
  for (size_t i = 0; i < a.size(); ++i)
  {
    put_u64(a.names[i].size());
    out += a.names[i];
    put_u64(a.sites);
    const uint8_t * r = a.row(i);
    for (size_t k = 0; k < a.sites; k += 2)
      out.push_back((char) ((code[r[k]] << 4) | (k + 1 < a.sites ? code[r[k + 1]] : 0)));
  }
  FILE * fh = std::fopen(out_path.c_str(), "wb");
  if (!fh) throw std::runtime_error("Cannot open file for writing: " + out_path);
  const bool ok = std::fwrite(out.data(), 1, out.size(), fh) == out.size();
  std::fclose(fh);
  if (!ok) throw std::runtime_error("Cannot write file: " + out_path);
  return out_path;
}

Alignment read_alignment(const std::string & path)
{
  return is_bfast(path) ? read_bfast(path) : read_fasta(path);
}

Alignment read_fasta(const std::string & path)
{
  FILE * fh = std::fopen(path.c_str(), "rb");
  if (!fh) throw std::runtime_error("Cannot open file: " + path);
  std::fseek(fh, 0, SEEK_END);
  const long size = std::ftell(fh);
  std::fseek(fh, 0, SEEK_SET);
  std::string text((size_t) (size > 0 ? size : 0), '\0');
  if (size > 0 && std::fread(&text[0], 1, (size_t) size, fh) != (size_t) size)
  {
    std::fclose(fh);
    throw std::runtime_error("Cannot read file: " + path);
  }
  std::fclose(fh);

  Alignment a;
  size_t pos = 0;
  size_t cur_len = 0;
  bool in_seq = false;
  auto finish = [&]() {
    if (!in_seq) return;
    if (a.names.size() == 1) a.sites = cur_len;
    else if (cur_len != a.sites)
      throw std::runtime_error(path + " does not contain equal size sequences! First offending sequence: " + a.names.back());
  };
  while (pos < text.size())
  {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    size_t end = eol;
    while (end > pos && (text[end - 1] == '\r' || text[end - 1] == ' ' || text[end - 1] == '\t')) --end;
    if (end > pos)
    {
      if (text[pos] == '>')
      {
        finish();
        a.names.emplace_back(text.substr(pos + 1, end - pos - 1));
        in_seq = true;
        cur_len = 0;
      }
      else
      {
        if (!in_seq) throw std::runtime_error(path + ": sequence data before the first '>' line");
        for (size_t i = pos; i < end; ++i)
        {
          const unsigned char c = (unsigned char) text[i];
          if (std::isspace(c)) continue;
          a.rows.push_back((uint8_t) std::toupper(c));
          ++cur_len;
        }
      }
    }
    pos = eol + 1;
  }
  finish();
  if (a.names.empty()) throw std::runtime_error(path + " contains no sequences");
  return a;
}

std::vector<uint8_t> gap_mask(const Alignment & a)
{
  bool is_gap[256] = {};
  for (const char * p = "NOX.-?nox"; *p; ++p) is_gap[(unsigned char) *p] = true;
  std::vector<uint8_t> mask(a.sites, 1);
  for (size_t i = 0; i < a.size(); ++i)
  {
    const uint8_t * r = a.row(i);
    for (size_t s = 0; s < a.sites; ++s) mask[s] &= (uint8_t) is_gap[r[s]];
  }
  return mask;
}

Alignment apply_mask(const Alignment & a, const std::vector<uint8_t> & drop)
{
  Alignment out;
  out.names = a.names;
  std::vector<size_t> keep;
  for (size_t s = 0; s < a.sites; ++s) if (!drop[s]) keep.push_back(s);
  out.sites = keep.size();
  out.rows.resize(a.size() * out.sites);
  for (size_t i = 0; i < a.size(); ++i)
  {
    const uint8_t * r = a.row(i);
    uint8_t * w = out.rows.data() + i * out.sites;
    for (size_t k = 0; k < keep.size(); ++k) w[k] = r[keep[k]];
  }
  return out;
}

}  // namespace epa_host
