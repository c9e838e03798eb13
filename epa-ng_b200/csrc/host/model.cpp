#include "model.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <fstream>
#include <iterator>
#include <cstdio>
#include <sstream>
#include <stdexcept>

namespace epa_host {

// ----------------------------------------------------------------------------------------------
//  Discrete GAMMA (Yang 1994). The building blocks are the classic published routines:
//  ln Gamma by Pike & Hill (CACM algorithm 291), the incomplete gamma ratio by Bhattacharjee
//  (AS 32), the normal quantile by Odeh & Evans (AS 70) and the chi-square quantile by Best &
//  Roberts (AS 91). libpll uses the same ones, so the category rates agree to rounding.
// ----------------------------------------------------------------------------------------------
namespace {

double ln_gamma(double a)
{
  double x = a, corr = 0.0;
  if (x < 7.0)
  {
    double f = 1.0, z = x;
    for (; z < 7.0; z += 1.0) f *= z;
    x = z;
    corr = -std::log(f);
  }
  const double z = 1.0 / (x * x);
  const double series = (((-.000595238095238 * z + .000793650793651) * z - .002777777777778) * z + .083333333333333) / x;
  return corr + (x - 0.5) * std::log(x) - x + .918938533204673 + series;
}

double incomplete_gamma_ratio(double x, double alpha, double ln_gamma_alpha)
{
  constexpr double accurate = 1e-8, overflow = 1e30;
  if (x == 0.0) return 0.0;
  if (x < 0.0 || alpha <= 0.0) return -1.0;
  const double factor = std::exp(alpha * std::log(x) - x - ln_gamma_alpha);
  if (x <= 1.0 || x < alpha)
  {
    // series expansion
    double gin = 1.0, term = 1.0, rn = alpha;
    do { rn += 1.0; term *= x / rn; gin += term; } while (term > accurate);
    return gin * factor / alpha;
  }
  // continued fraction
  double a = 1.0 - alpha, b = a + x + 1.0, term = 0.0;
  double pn[6] = {1.0, x, x + 1.0, x * b, 0.0, 0.0};
  double gin = pn[2] / pn[3];
  for (;;)
  {
    a += 1.0; b += 2.0; term += 1.0;
    const double an = a * term;
    pn[4] = b * pn[2] - an * pn[0];
    pn[5] = b * pn[3] - an * pn[1];
    if (pn[5] != 0.0)
    {
      const double rn = pn[4] / pn[5];
      const double dif = std::fabs(gin - rn);
      if (dif <= accurate && dif <= accurate * rn) return 1.0 - factor * gin;
      gin = rn;
    }
    for (int i = 0; i < 4; ++i) pn[i] = pn[i + 2];
    if (std::fabs(pn[4]) >= overflow)
      for (int i = 0; i < 4; ++i) pn[i] /= overflow;
  }
}

double normal_quantile(double prob)
{
  constexpr double a0 = -.322232431088, a1 = -1.0, a2 = -.342242088547, a3 = -.0204231210245, a4 = -.453642210148e-4;
  constexpr double b0 = .0993484626060, b1 = .588581570495, b2 = .531103462366, b3 = .103537752850, b4 = .0038560700634;
  const double p1 = prob < 0.5 ? prob : 1.0 - prob;
  if (p1 < 1e-20) return -9999.0;
  const double y = std::sqrt(std::log(1.0 / (p1 * p1)));
  const double z = y + ((((y * a4 + a3) * y + a2) * y + a1) * y + a0) / ((((y * b4 + b3) * y + b2) * y + b1) * y + b0);
  return prob < 0.5 ? -z : z;
}

double chi2_quantile(double prob, double v)
{
  constexpr double e = .5e-6, aa = .6931471805;
  const double p = prob;
  if (p < .000002 || p > .999998 || v <= 0.0) return -1.0;
  const double g = ln_gamma(v / 2.0);
  const double xx = v / 2.0, c = xx - 1.0;
  double ch;
  if (v < -1.24 * std::log(p))
  {
    ch = std::pow(p * xx * std::exp(g + xx * aa), 1.0 / xx);
    if (ch - e < 0.0) return ch;
  }
  else if (v <= .32)
  {
    ch = 0.4;
    const double a = std::log(1.0 - p);
    double q;
    do
    {
      q = ch;
      const double p1 = 1.0 + ch * (4.67 + ch);
      const double p2 = ch * (6.73 + ch * (6.66 + ch));
      const double t = -0.5 + (4.67 + 2.0 * ch) / p1 - (6.73 + ch * (13.32 + 3.0 * ch)) / p2;
      ch -= (1.0 - std::exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
    } while (std::fabs(q / ch - 1.0) - .01 > 0.0);
  }
  else
  {
    const double x = normal_quantile(p);
    const double p1 = 0.222222 / v;
    ch = v * std::pow(x * std::sqrt(p1) + 1.0 - p1, 3.0);
    if (ch > 2.2 * v + 6.0) ch = -2.0 * (std::log(1.0 - p) - c * std::log(.5 * ch) + g);
  }
  double q;
  do
  {
    q = ch;
    const double p1 = .5 * ch;
    double t = incomplete_gamma_ratio(p1, xx, g);
    if (t < 0.0) return -1.0;
    const double p2 = p - t;
    t = p2 * std::exp(xx * aa + g + p1 - c * std::log(ch));
    const double b = t / ch, a = 0.5 * t - b * c;
    const double s1 = (210.0 + a * (140.0 + a * (105.0 + a * (84.0 + a * (70.0 + 60.0 * a))))) / 420.0;
    const double s2 = (420.0 + a * (735.0 + a * (966.0 + a * (1141.0 + 1278.0 * a)))) / 2520.0;
    const double s3 = (210.0 + a * (462.0 + a * (707.0 + 932.0 * a))) / 2520.0;
    const double s4 = (252.0 + a * (672.0 + 1182.0 * a) + c * (294.0 + a * (889.0 + 1740.0 * a))) / 5040.0;
    const double s5 = (84.0 + 264.0 * a + c * (175.0 + 606.0 * a)) / 2520.0;
    const double s6 = (120.0 + c * (346.0 + 127.0 * c)) / 5040.0;
    ch += t * (1.0 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
  } while (std::fabs(q / ch - 1.0) > e);
  return ch;
}

}  // namespace

std::vector<double> discrete_gamma_rates(double alpha, int ncat, bool median)
{
  if (alpha < 0.02) throw std::runtime_error("GAMMA alpha must be >= 0.02");
  std::vector<double> out((size_t) ncat, 1.0);
  if (ncat == 1) return out;
  const double factor = alpha / alpha * ncat, beta = alpha;
  if (median)
  {
    const double middle = 1.0 / (2.0 * ncat);
    double t = 0.0;
    for (int i = 0; i < ncat; ++i)
    {
      out[i] = chi2_quantile((i * 2.0 + 1.0) * middle, 2.0 * alpha) / (2.0 * beta);
      t += out[i];
    }
    for (int i = 0; i < ncat; ++i) out[i] *= factor / t;
    return out;
  }
  std::vector<double> cut((size_t) ncat - 1);
  const double lg = ln_gamma(alpha + 1.0);
  for (int i = 0; i < ncat - 1; ++i) cut[i] = chi2_quantile((i + 1.0) / ncat, 2.0 * alpha) / (2.0 * beta);
  for (int i = 0; i < ncat - 1; ++i) cut[i] = incomplete_gamma_ratio(cut[i] * beta, alpha + 1.0, lg);
  out[0] = cut[0] * factor;
  out[ncat - 1] = (1.0 - cut[ncat - 2]) * factor;
  for (int i = 1; i < ncat - 1; ++i) out[i] = (cut[i] - cut[i - 1]) * factor;
  return out;
}

// ----------------------------------------------------------------------------------------------
//  Eigen system. The reference tridiagonalises (Householder) and runs QL; a cyclic Jacobi sweep
//  gives the same decomposition up to ordering and rotations inside degenerate eigenspaces, to
//  which P(t) and the sumtable contractions are invariant.
// ----------------------------------------------------------------------------------------------
void eigen_decompose(int S, const std::vector<double> & subst, const std::vector<double> & freqs,
                     std::vector<double> & eigenvals, std::vector<double> & eigenvecs,
                     std::vector<double> & inv_eigenvecs)
{
  std::vector<double> a((size_t) S * S, 0.0), v((size_t) S * S, 0.0);
  auto A = [&](int i, int j) -> double & { return a[(size_t) i * S + j]; };
  auto V = [&](int i, int j) -> double & { return v[(size_t) i * S + j]; };
  const double last = subst.back();
  int k = 0;
  for (int i = 0; i < S; ++i)
    for (int j = i + 1; j < S; ++j)
    {
      double r = subst[k++];
      if (last > 0.0) r /= last;
      A(i, j) = A(j, i) = r * std::sqrt(freqs[i] * freqs[j]);
      A(i, i) -= r * freqs[j];
      A(j, j) -= r * freqs[i];
    }
  double mean = 0.0;
  for (int i = 0; i < S; ++i) mean += freqs[i] * (-A(i, i));
  for (double & x : a) x /= mean;
  for (int i = 0; i < S; ++i) V(i, i) = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep)
  {
    double off = 0.0;
    for (int i = 0; i < S; ++i) for (int j = i + 1; j < S; ++j) off += A(i, j) * A(i, j);
    if (off < 1e-300) break;
    for (int p = 0; p < S; ++p)
      for (int q = p + 1; q < S; ++q)
      {
        if (A(p, q) == 0.0) continue;
        const double theta = (A(q, q) - A(p, p)) / (2.0 * A(p, q));
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int r = 0; r < S; ++r) { const double x = A(r, p), y = A(r, q); A(r, p) = c * x - s * y; A(r, q) = s * x + c * y; }
        for (int r = 0; r < S; ++r) { const double x = A(p, r), y = A(q, r); A(p, r) = c * x - s * y; A(q, r) = s * x + c * y; }
        for (int r = 0; r < S; ++r) { const double x = V(r, p), y = V(r, q); V(r, p) = c * x - s * y; V(r, q) = s * x + c * y; }
      }
  }
  eigenvals.assign(S, 0.0);
  eigenvecs.assign((size_t) S * S, 0.0);
  inv_eigenvecs.assign((size_t) S * S, 0.0);
  for (int m = 0; m < S; ++m)
  {
    eigenvals[m] = A(m, m);
    for (int j = 0; j < S; ++j)
    {
      eigenvecs[(size_t) m * S + j] = V(j, m) * std::sqrt(freqs[j]);
      inv_eigenvecs[(size_t) j * S + m] = V(j, m) / std::sqrt(freqs[j]);
    }
  }
}

// ----------------------------------------------------------------------------------------------
//  Model string
// ----------------------------------------------------------------------------------------------
namespace {

std::vector<double> braces(const std::string & s, size_t & pos, bool & present)
{
  std::vector<double> out;
  present = false;
  if (pos < s.size() && s[pos] == '{')
  {
    const size_t end = s.find('}', pos);
    if (end == std::string::npos) throw std::runtime_error("model: missing '}'");
    std::stringstream ss(s.substr(pos + 1, end - pos - 1));
    std::string tok;
    while (std::getline(ss, tok, '/')) out.push_back(std::stod(tok));
    pos = end + 1;
    present = true;
  }
  return out;
}

}  // namespace

Model Model::parse(const std::string & desc)
{
  Model m;
  size_t pos = desc.find_first_of("+{[");
  if (pos == std::string::npos) pos = desc.size();
  std::string name = desc.substr(0, pos);
  std::transform(name.begin(), name.end(), name.begin(), [](unsigned char c) { return (char) std::toupper(c); });
  std::string opts = desc.substr(pos);
  if (name == "DNA") { name = "GTR"; opts = "+G+FO"; }

  // rate symmetries AC AG AT CG CT GT of the named DNA models and their aliases (PM/util/models_dna.c:40-125)
  static const struct { const char * name; int sym[6]; } kDna[] = {
    {"JC", {0, 0, 0, 0, 0, 0}}, {"F81", {0, 0, 0, 0, 0, 0}}, {"K80", {0, 1, 0, 0, 1, 0}}, {"HKY", {0, 1, 0, 0, 1, 0}},
    {"TN93EF", {0, 1, 0, 0, 2, 0}}, {"TN93", {0, 1, 0, 0, 2, 0}}, {"TRNEF", {0, 1, 0, 0, 2, 0}}, {"TRN", {0, 1, 0, 0, 2, 0}},
    {"K81", {0, 1, 2, 2, 1, 0}}, {"K81UF", {0, 1, 2, 2, 1, 0}}, {"TPM1", {0, 1, 2, 2, 1, 0}}, {"TPM1UF", {0, 1, 2, 2, 1, 0}},
    {"TPM2", {0, 1, 0, 2, 1, 2}}, {"TPM2UF", {0, 1, 0, 2, 1, 2}}, {"TPM2EF", {0, 1, 0, 2, 1, 2}},
    {"TPM3", {0, 1, 2, 0, 1, 2}}, {"TPM3UF", {0, 1, 2, 0, 1, 2}}, {"TPM3EF", {0, 1, 2, 0, 1, 2}},
    {"TIM1", {0, 1, 2, 2, 3, 0}}, {"TIM1UF", {0, 1, 2, 2, 3, 0}}, {"TIM1EF", {0, 1, 2, 2, 3, 0}},
    {"TIM2", {0, 1, 0, 2, 3, 2}}, {"TIM2UF", {0, 1, 0, 2, 3, 2}}, {"TIM2EF", {0, 1, 0, 2, 3, 2}},
    {"TIM3", {0, 1, 2, 0, 3, 2}}, {"TIM3UF", {0, 1, 2, 0, 3, 2}}, {"TIM3EF", {0, 1, 2, 0, 3, 2}},
    {"TVMEF", {0, 1, 2, 3, 1, 4}}, {"TVM", {0, 1, 2, 3, 1, 4}}, {"SYM", {0, 1, 2, 3, 4, 5}}, {"GTR", {0, 1, 2, 3, 4, 5}}};
  std::vector<int> sym;
  bool dna = false;
  for (const auto & d : kDna)
    if (name == d.name) { sym.assign(d.sym, d.sym + 6); dna = true; break; }

  if (dna)
  {
    m.states = 4;
    m.freqs.assign(4, 0.25);
    // JC / F81 carry the model's equal rates; every other DNA model without {rates} starts from the ML-mode default
    // 0.5 0.5 0.5 0.5 0.5 1.0 over the SIX rates, whatever its symmetry (src/core/raxml/Model.cpp:484-490)
    if (name == "JC" || name == "F81") m.subst.assign(6, 1.0);
    else m.subst = {0.5, 0.5, 0.5, 0.5, 0.5, 1.0};
  }
  else if (name == "PROTGTR")
  {
    // protein GTR: 190 user exchangeabilities in the order of the upper triangle (PM/util/models_aa.c:69); ML-mode defaults otherwise
    m.states = 20;
    m.subst.assign(190, 0.5);
    m.subst.back() = 1.0;
    m.freqs.assign(20, 1.0 / 20);
  }
  else
  {
    m.states = 20;
    if (!protein_model(name, m.subst, m.freqs)) throw std::runtime_error("unsupported model name: " + name);
  }
  m.name = name;

  size_t i = 0;
  bool present = false;
  std::vector<double> vals = braces(opts, i, present);
  if (present)
  {
    if (!dna && name != "PROTGTR") throw std::runtime_error("user-defined rates need PROTGTR for protein data");
    if (!dna)
    {
      if (vals.size() != 190) throw std::runtime_error("model: wrong number of substitution rates");
      for (size_t k = 0; k < 190; ++k) m.subst[k] = vals[k] / vals.back();
    }
    else
    {
      const int nuniq = *std::max_element(sym.begin(), sym.end()) + 1;
      if ((int) vals.size() != nuniq) throw std::runtime_error("model: wrong number of substitution rates");
      const double last = vals[sym.back()];
      for (int k = 0; k < 6; ++k) m.subst[k] = vals[sym[k]] / last;
    }
  }
  bool gamma = false;
  std::vector<double> free_rates, free_weights;
  while (i < opts.size())
  {
    const char ch = (char) std::toupper((unsigned char) opts[i++]);
    if (ch == '+') continue;
    if (ch == 'F')
    {
      char mode = 'C';
      if (i < opts.size() && opts[i] != '+' && opts[i] != '{') mode = (char) std::toupper((unsigned char) opts[i++]);
      if (mode == 'U')
      {
        vals = braces(opts, i, present);
        if ((int) vals.size() != m.states) throw std::runtime_error("model: wrong number of base frequencies");
        double sum = 0.0;
        for (double v : vals) sum += v;
        for (int k = 0; k < m.states; ++k) m.freqs[k] = vals[k] / sum;
      }
      else if (mode == 'E' || mode == 'O') m.freqs.assign(m.states, 1.0 / m.states);
      else if (mode == 'C') m.empirical_freqs = true;      // counted once the reference MSA is linked
      else throw std::runtime_error("Invalid frequencies specification: " + desc);
    }
    else if (ch == 'G')
    {
      gamma = true;
      std::string num;
      while (i < opts.size() && std::isdigit((unsigned char) opts[i])) num += opts[i++];
      m.rate_cats = num.empty() ? 4 : std::stoi(num);
      if (i < opts.size() && (opts[i] == 'a' || opts[i] == 'A')) { m.gamma_median = true; ++i; }
      else if (i < opts.size() && (opts[i] == 'm' || opts[i] == 'M')) ++i;
      vals = braces(opts, i, present);
      if (present) m.alpha = vals.at(0);
    }
    else if (ch == 'R')
    {
      // free rates (src/core/raxml/Model.cpp:405-455): +R[n]{rates}{weights}; weights normalised to sum 1,
      // rates to mean 1; without values the categories start as GAMMA(alpha = 1) with equal weights
      gamma = true;
      std::string num;
      while (i < opts.size() && std::isdigit((unsigned char) opts[i])) num += opts[i++];
      if (!num.empty()) m.rate_cats = std::stoi(num); else if (m.rate_cats == 1) m.rate_cats = 4;
      vals = braces(opts, i, present);
      if (present)
      {
        if ((int) vals.size() != m.rate_cats) throw std::runtime_error("Invalid number of free rates specified: " + desc);
        free_rates = vals;
        vals = braces(opts, i, present);
        if (present)
        {
          if ((int) vals.size() != m.rate_cats) throw std::runtime_error("Invalid number of rate weights specified: " + desc);
          double sum = 0.0;
          for (double w : vals) sum += w;
          for (double & w : vals) w /= sum;
          free_weights = vals;
        }
        else free_weights.assign((size_t) m.rate_cats, 1.0 / m.rate_cats);
        double mean = 0.0;
        for (size_t k = 0; k < free_rates.size(); ++k) mean += free_rates[k] * free_weights[k];
        for (double & r : free_rates) r /= mean;
      }
    }
    else if (ch == 'I')
    {
      // src/core/raxml/Model.cpp:355-380: +I / +IO = ML mode (stays at the reference's unoptimised start
      // value 0, Model.cpp:192), +IU{p} = user value, +IC = empirical (needs alignment statistics)
      char mode = 'O';
      if (i < opts.size() && opts[i] != '+') mode = (char) std::toupper((unsigned char) opts[i++]);
      if (mode == 'U')
      {
        vals = braces(opts, i, present);
        if (!present || vals.empty()) throw std::runtime_error("Invalid p-inv specification: " + desc);
        m.pinv = vals[0];
        if (!(m.pinv >= 0.0 && m.pinv < 1.0)) throw std::runtime_error("Invalid proportion of invariant sites: " + desc);
      }
      else if (mode == 'C') m.pinv = 0.0;     // the reference never counts the empirical value ("P-inv (empirical): 0")
      else if (mode != 'O')
        throw std::runtime_error("Invalid p-inv specification: " + desc);
    }
    else
      throw std::runtime_error(std::string("unsupported model option +") + ch + " (supported: +F{U,E,O,C}, +G, +R, +I{U,O,C})");
  }
  m.rates = (gamma && m.rate_cats > 1) ? discrete_gamma_rates(m.alpha, m.rate_cats, m.gamma_median)
                                       : std::vector<double>((size_t) m.rate_cats, 1.0);
  m.weights.assign((size_t) m.rate_cats, 1.0 / m.rate_cats);
  if (!free_rates.empty()) { m.rates = free_rates; m.weights = free_weights; }
  {
    // pll_set_frequencies (LP/models.c:445-470): frequencies that do not sum to 1 within 1e-8 are normalised
    // (the published protein tables carry six digits: LG sums to 1.000001, WAG to 0.9999999)
    double sum = 0.0;
    for (double f : m.freqs) sum += f;
    if (std::fabs(sum - 1.0) > 1e-8)
      for (double & f : m.freqs) f /= sum;
  }
  eigen_decompose(m.states, m.subst, m.freqs, m.eigenvals, m.eigenvecs, m.inv_eigenvecs);
  return m;
}

void Model::set_empirical_freqs(const uint32_t * tip_masks, size_t n_tips, size_t sites)
{
  std::vector<double> f((size_t) states, 0.0);
  for (size_t i = 0; i < n_tips * sites; ++i)
  {
    uint32_t st = tip_masks[i];
    const double share = 1.0 / (double) __builtin_popcount(st);
    for (int k = 0; k < states; ++k, st >>= 1)
      if (st & 1u) f[(size_t) k] += share;
  }
  for (double & v : f) v /= (double) (n_tips * sites);
  freqs = f;
  eigen_decompose(states, subst, freqs, eigenvals, eigenvecs, inv_eigenvecs);
}

// ---- model files (-m <file>, src/main.cpp:433-436 -> src/util/parse_model.hpp) ------------------------------
namespace {

// text after `key` (searched from pos) up to the end of its line; pos moves to that line end
std::string field_after(const std::string & text, const std::string & key, size_t & pos)
{
  const size_t at = text.find(key, pos);
  if (at == std::string::npos) throw std::invalid_argument("Couldn't parse model file! (can't find '" + key + "'!)");
  const size_t from = at + key.size(), eol = text.find('\n', from);
  if (eol == std::string::npos) throw std::runtime_error("couldnt find terminating newline?!");
  pos = eol;
  return text.substr(from, eol - from);
}

bool later_has(const std::string & text, const std::string & key, size_t pos) { return text.find(key, pos) != std::string::npos; }

// "{r01/r02/.../r(S-2)(S-1)}+FU{f0/.../f(S-1)}" collected pair by pair in file order
std::string rates_and_freqs(const std::string & text, size_t & pos, const std::string & states, const std::string & rate_prefix,
                            const std::string & rate_infix, const std::string & rate_suffix, const std::string & freq_prefix,
                            const std::string & freq_suffix)
{
  std::string out = "{";
  for (size_t i = 0; i + 1 < states.size(); ++i)
    for (size_t k = i + 1; k < states.size(); ++k)
    {
      if (k > 1) out += "/";
      out += field_after(text, rate_prefix + states[i] + rate_infix + states[k] + rate_suffix, pos);
    }
  out += "}+FU{";
  for (size_t i = 0; i < states.size(); ++i)
  {
    if (i) out += "/";
    out += field_after(text, freq_prefix + states[i] + freq_suffix, pos);
  }
  return out + "}";
}

}  // namespace

std::string model_string_from_file(const std::string & path)
{
  std::ifstream in(path);
  if (!in) throw std::runtime_error("Cannot open model file: " + path);
  const std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const std::string first_line = text.substr(0, text.find('\n'));
  const char * dna_states = "ACGT", * aa_states = "ARNDCQEGHILKMFPSTWYV";
  size_t pos = 0;
  if (first_line.rfind("IQ-TREE ", 0) == 0)
  {
    // IQ-TREE report: "Model of substitution: GTR+F+I+G4", "A-C: ..", "pi(A) = ..", "Gamma with 4 categories", ...
    const std::string iq = field_after(text, "Model of substitution: ", pos);
    const std::string matrix = iq.substr(0, iq.find('+'));
    std::string desc = matrix + rates_and_freqs(text, pos, matrix == "GTR" ? dna_states : aa_states, "", "-", ": ", "pi(", ") = ");
    std::string cats;
    const bool gamma = later_has(text, "Gamma with ", pos);
    if (gamma)
    {
      const std::string tail = field_after(text, "Gamma with ", pos);
      const size_t end = tail.find(" categories");
      if (end == std::string::npos) throw std::invalid_argument("Couldn't parse model file! (can't find ' categories'!)");
      if (end == 0) throw std::runtime_error("Nothing inbetween ' categories' and 'Gamma with '?");
      cats = tail.substr(0, end);
    }
    if (later_has(text, "Proportion of invariable sites: ", pos)) desc += "+IU{" + field_after(text, "Proportion of invariable sites: ", pos) + "}";
    if (gamma) desc += "+G" + cats + "{" + field_after(text, "Gamma shape alpha: ", pos) + "}";
    return desc;
  }
  if (text.find("This is RAxML version 8.") != std::string::npos)
  {
    // RAxML 8 info file (-f e): DataType, Substitution Matrix, alpha, invar, "rate A <-> C: ..", "freq pi(A): .."
    const bool dna = field_after(text, "DataType: ", pos) == "DNA";
    std::string matrix = field_after(text, "Substitution Matrix: ", pos);
    if (!dna && matrix == "GTR") matrix = "PROTGTR";
    std::string alpha, pinv;
    if (later_has(text, "alpha: ", pos)) alpha = "+G4{" + field_after(text, "alpha: ", pos) + "}";
    if (later_has(text, "invar: ", pos)) pinv = "+IU{" + field_after(text, "invar: ", pos) + "}";
    return matrix + rates_and_freqs(text, pos, dna ? dna_states : aa_states, "rate ", " <-> ", ": ", "freq pi(", "): ") + pinv + alpha;
  }
  // raxml-ng .bestModel: "<model string>, <partition> = <range>"
  const size_t comma = first_line.find(',');
  if (comma == std::string::npos) throw std::runtime_error("Model string in provided file seems wrong.");
  return first_line.substr(0, comma);
}

std::string Model::describe() const
{
  std::ostringstream os;
  os << "   Rate heterogeneity: ";
  if (rate_cats > 1)
  {
    os << "GAMMA (" << rate_cats << " cats, " << (gamma_median ? "median" : "mean") << "),  alpha: " << alpha
       << " (user),  weights&rates: ";
    for (int r = 0; r < rate_cats; ++r) os << "(" << weights[r] << "," << rates[r] << ") ";
  }
  else os << "NONE";
  if (pinv > 0.0) os << "\n        P-inv (user): " << pinv;
  os << "\n        Base frequencies (" << (empirical_freqs ? "empirical" : "user") << "): ";
  for (double f : freqs) os << f << " ";
  if (states == 4)
  {
    os << "\n        Substitution rates (user): ";
    for (double r : subst) os << r << " ";
  }
  else os << "\n        Substitution matrix: " << name;
  return os.str();
}

}  // namespace epa_host
