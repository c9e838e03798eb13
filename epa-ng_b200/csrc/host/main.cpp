// main.cpp - command line of the B200 build. Accepts the reference's flag spellings
// (/root/reference/src/main.cpp:96-261); flags that only make sense for the CPU program
// (-T/--threads) are accepted and ignored, modes outside the hot path are rejected by name.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "epa_b200_host.h"

namespace {

[[noreturn]] void die(const std::string & msg)
{
  std::fprintf(stderr, "ERROR %s\n", msg.c_str());
  std::exit(1);
}

void usage()
{
  std::puts(
      "epa-ng-b200 - massively parallel phylogenetic placement of genetic sequences (B200 build)\n"
      "Usage: epa-ng-b200 -t TREE -s REF_MSA -q QUERY -m MODEL [-w OUTDIR] [options]\n"
      "  -t,--tree FILE            reference tree (newick, unrooted binary)\n"
      "  -s,--ref-msa,--msa FILE   reference alignment (fasta)\n"
      "  -q,--query FILE           aligned query sequences (fasta)\n"
      "  -m,--model STRING         e.g. GTR{0.7/1.8/1.2/0.6/3.0/1.0}+FU{0.25/0.23/0.30/0.22}+G4{0.47}\n"
      "  -w,--outdir DIR           output directory [./]\n"
      "  -g,--dyn-heur FLOAT       accumulated-LWR preplacement heuristic [0.99999]\n"
      "  -G,--fix-heur FLOAT       fixed-fraction preplacement heuristic\n"
      "  --baseball-heur           baseball heuristic (strike box 3, 6 strikes, 40 pitches)\n"
      "  --no-heur                 evaluate every edge thoroughly\n"
      "  --raxml-blo               optimise the three branch lengths the way RAxML-EPA did (default: pplacer-style)\n"
      "  --chunk-size INT          queries per device chunk [131072]\n"
      "  --no-pre-mask             do not drop all-gap columns / trim query ranges\n"
      "  --filter-acc-lwr FLOAT    accumulated-LWR output filter\n"
      "  --filter-min-lwr FLOAT    minimum-LWR output filter [0.01]\n"
      "  --filter-min INT [1]      --filter-max INT [7]      --precision INT [10]\n"
      "  -c,--bfast FILE           convert an aligned DNA FASTA file to the bfast format (into -w) and exit\n"
      "  --device INT              CUDA device [0]\n"
      "  --devices LIST            several GPUs of this box, e.g. 0,1,2,3 or 0-7: query chunks are handed out\n"
      "                            in file order, one host thread per GPU, one shared reader and jplace writer\n"
      "  -T, --threads N           host threads for reading, decoding and formatting [all]\n"
      "  --redo, --verbose         accepted for compatibility\n"
      "  -v,--version\n");
}

}  // namespace

int main(int argc, char ** argv)
{
  std::string tree, ref, query, model = "GTR+G", outdir = "./", bfast_conv;
  epa_options opts;
  epa_options_default(&opts);
  uint32_t chunk = 0;
  int precision = 10, device = 0, preserve_rooting = 1, rate_mode = 2, rate_bug = 1, host_threads = 0;
  std::vector<int> devices;
  std::string invocation;
  for (int i = 0; i < argc; ++i) { invocation += argv[i]; invocation += ' '; }

  auto need = [&](int & i) -> const char * {
    if (i + 1 >= argc) die(std::string("option ") + argv[i] + " needs a value");
    return argv[++i];
  };
  for (int i = 1; i < argc; ++i)
  {
    const std::string a = argv[i];
    if (a == "-h" || a == "--help") { usage(); return 0; }
    else if (a == "-v" || a == "--version") { std::puts("EPA-ng-b200 v0.1 (placement hot path of EPA-ng v0.3.8 on sm_100a)"); return 0; }
    else if (a == "-t" || a == "--tree") tree = need(i);
    else if (a == "-s" || a == "--ref-msa" || a == "--msa") ref = need(i);
    else if (a == "-q" || a == "--query") query = need(i);
    else if (a == "-m" || a == "--model") model = need(i);
    else if (a == "-w" || a == "--outdir") outdir = need(i);
    else if (a == "-g" || a == "--dyn-heur") { opts.prescoring = 1; opts.heuristic = 0; opts.prescoring_threshold = std::atof(need(i)); }
    else if (a == "-G" || a == "--fix-heur") { opts.prescoring = 1; opts.heuristic = 1; opts.prescoring_threshold = std::atof(need(i)); }
    else if (a == "--baseball-heur") { opts.prescoring = 1; opts.heuristic = 2; }
    else if (a == "--no-heur") opts.prescoring = 0;
    else if (a == "--chunk-size") chunk = (uint32_t) std::atol(need(i));
    else if (a == "--no-pre-mask") opts.premasking = 0;
    else if (a == "--filter-acc-lwr") { opts.filter_acc_lwr = 1; opts.support_threshold = std::atof(need(i)); }
    else if (a == "--filter-min-lwr") { opts.filter_acc_lwr = 0; opts.support_threshold = std::atof(need(i)); }
    else if (a == "--filter-min") opts.filter_min = (uint32_t) std::atol(need(i));
    else if (a == "--filter-max") opts.filter_max = (uint32_t) std::atol(need(i));
    else if (a == "--precision") precision = std::atoi(need(i));
    else if (a == "--device") device = std::atoi(need(i));
    else if (a == "--devices")
    {
      // "0,1,2" and ranges "0-7"
      const std::string list = need(i);
      devices.clear();
      size_t pos = 0;
      while (pos < list.size())
      {
        size_t end = list.find(',', pos);
        if (end == std::string::npos) end = list.size();
        const std::string item = list.substr(pos, end - pos);
        const size_t dash = item.find('-');
        if (dash != std::string::npos && dash > 0)
          for (int d = std::atoi(item.substr(0, dash).c_str()); d <= std::atoi(item.substr(dash + 1).c_str()); ++d) devices.push_back(d);
        else if (!item.empty()) devices.push_back(std::atoi(item.c_str()));
        pos = end + 1;
      }
    }
    else if (a == "-T" || a == "--threads") host_threads = std::atoi(need(i));
    else if (a == "--tmp") (void) need(i);
    else if (a == "--redo" || a == "--verbose") {}
    else if (a == "--rate-scalers")
    {
      const std::string v = need(i);
      if (v == "on") rate_mode = 1; else if (v == "off") rate_mode = 0; else if (v == "auto") rate_mode = 2;
      else die("--rate-scalers takes on, off or auto");
    }
    else if (a == "--correct-scaler-focus") rate_bug = 0;     // not a reference option: read the scalers of the site itself
    else if (a == "--preserve-rooting") { const std::string v = need(i); preserve_rooting = (v != "off"); }
    else if (a == "--raxml-blo") opts.sliding_blo = 0;        // src/core/pll/optimize.cpp:274-278
    else if (a == "-c" || a == "--bfast") bfast_conv = need(i);        // src/main.cpp:284-288
    else if (a == "-b" || a == "--binary" || a == "-B" || a == "--dump-binary" || a == "--split")
      die("option " + a + " is outside the accelerated hot path and not supported by this build");
    else die("unknown option " + a);
  }
  if (!bfast_conv.empty())
  {
    char written[4096];
    if (epa_host_fasta_to_bfast(bfast_conv.c_str(), outdir.c_str(), written, sizeof written)) die(epa_host_last_error());
    std::printf("Resulting bfast file was written to: %s\n", written);
    return 0;
  }
  if (tree.empty() || ref.empty() || query.empty()) { usage(); die("-t, -s and -q are required"); }
  if (epa_host_set_rate_scalers(rate_mode, rate_bug)) die(epa_host_last_error());
  if (devices.empty()) devices.push_back(device);
  const int rc = epa_run_files_multi(tree.c_str(), ref.c_str(), query.c_str(), model.c_str(), outdir.c_str(), &opts, chunk,
                                     precision, devices.data(), (uint32_t) devices.size(), invocation.c_str(), preserve_rooting,
                                     host_threads, nullptr);
  if (rc) die(epa_host_last_error());
  return 0;
}
