#include "seqio.hpp"

#include <cctype>
#include <cstdio>
#include <stdexcept>

namespace epa_host {

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
