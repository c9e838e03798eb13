// protein_models.cpp - empirical amino-acid replacement matrices of the host layer.
// (No table is compiled in yet: protein reference trees can be served through the device API
// with caller-provided eigen systems; the host model parser reports the name as unsupported.)
#include "model.hpp"

namespace epa_host {

bool protein_model(const std::string &, std::vector<double> &, std::vector<double> &) { return false; }

}  // namespace epa_host
