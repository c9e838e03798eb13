// model.hpp - substitution model of the host layer: the subset of the raxml-ng model grammar that
// the reference accepts through -m (src/core/raxml/Model.cpp:123-560) and that the device path
// supports: the 22 named DNA models (JC ... GTR) with their aliases and the 28 empirical protein matrices of the reference (protein_models.cpp), with
// user, equal or empirical frequencies (+FU{..}/+FE/+FO, +F/+FC), discrete GAMMA rate heterogeneity (+G[n][a|m]{alpha}), free rates (+R[n]{rates}{weights}) and a
// user proportion of invariant sites (+IU{p}). Ascertainment correction is rejected with a clear message.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace epa_host {

struct Model {
  int states = 4;
  std::string name;
  std::vector<double> subst;        // upper triangle, S(S-1)/2, last entry = 1 after normalisation
  std::vector<double> freqs;        // S
  double alpha = 1.0;
  int rate_cats = 1;
  bool gamma_median = false;
  double pinv = 0.0;                // +IU{p} (src/core/raxml/Model.cpp:355-380); +I/+IO/+IC stay 0 as in the reference
  bool empirical_freqs = false;     // +F / +FC: frequencies are counted on the reference MSA (set_empirical_freqs)
  std::vector<double> rates, weights;
  std::vector<double> eigenvals, eigenvecs, inv_eigenvecs;   // libpll layout (models.c:394-404)

  static Model parse(const std::string & desc);   // throws std::runtime_error
  // compute_and_set_empirical_frequencies (src/core/pll/optimize.cpp:457-472 -> pllmod_msa_empirical_frequencies,
  // PM/msa/pll_msa.c:45-143): every tip character spreads one count evenly over the states of its mask; divided
  // by sites * tips. Recomputes the eigen system.
  void set_empirical_freqs(const uint32_t * tip_masks, size_t n_tips, size_t sites);
  std::string describe() const;                   // log text in the spirit of the reference's model print-out
};

// -m <file>: model string out of a RAxML 8 info file, a raxml-ng .bestModel file or an IQ-TREE report
// (src/main.cpp:433-436 -> src/util/parse_model.hpp); throws like the reference when a field is missing
std::string model_string_from_file(const std::string & path);

// Discrete GAMMA category rates, mean or median (Yang 1994; libpll gamma.c:220-292)
std::vector<double> discrete_gamma_rates(double alpha, int ncat, bool median);

// Eigen system of the symmetrised, mean-rate-normalised reversible rate matrix (libpll models.c:182-410)
void eigen_decompose(int S, const std::vector<double> & subst, const std::vector<double> & freqs,
                     std::vector<double> & eigenvals, std::vector<double> & eigenvecs,
                     std::vector<double> & inv_eigenvecs);

// Protein exchangeability tables (protein_models.cpp). Returns false when the name is unknown.
bool protein_model(const std::string & upper_name, std::vector<double> & subst, std::vector<double> & freqs);

}  // namespace epa_host
