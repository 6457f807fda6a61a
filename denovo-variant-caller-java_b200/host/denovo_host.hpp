// denovo_host.hpp — host side of the B200 trio de novo caller: the reference's
// plugin interface above the C ABI (include/denovo_b200.h).
//
// Mirrors, class for class, the part of alenzhao/denovo-variant-caller-java that sits
// around the hot path ("M/" = src/main/java/com/google/cloud/genomics/denovo/):
//   CommandLine      M/CommandLine.java:27-131   the 19 flags, names and defaults
//   DenovoShared     M/DenovoShared.java         immutable parameter bag
//   DenovoCaller     M/DenovoCaller.java:24-78   abstract execute(), parseLine, writeCalls
//   DenovoCallers    M/DenovoCallers.java:32-45  factory
//   ReadCaller       M/ReadCaller.java:57-130    candidate file loop -> GpuReadCaller (one batched call)
//   VariantCaller    M/VariantCaller.java:55-210 contig/stride shards -> GpuVariantCaller
//   DenovoRunner     M/DenovoRunner.java:165-186 mode dispatch, FULL chaining through <out>.tmp
// The Google Genomics API client (network; M/ReadCaller.java:155-186,
// M/VariantContigStream.java) is replaced by seeded synthetic sources.
// The reference is Java; this image has no JVM, so the compiled host is C++ and the
// Java (Panama FFM) binding ships as source under ../java.
#pragma once
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "denovo_b200.h"

namespace denovo {

// ---- enums (M/DenovoUtil.java:60,85-130,135-143,195-205,261-288) ---------------------------
enum class Caller { VARIANT, READ, FULL };
enum class LogLevel { ERROR, INFO, DEBUG };
static const char* const kGenotypeNames[10] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT"};
static const char kAlleleNames[4] = {'A', 'C', 'G', 'T'};
static const char* const kChromosomeNames[25] = {
    "CHR1",  "CHR2",  "CHR3",  "CHR4",  "CHR5",  "CHR6",  "CHR7",  "CHR8",  "CHR9",
    "CHR10", "CHR11", "CHR12", "CHR13", "CHR14", "CHR15", "CHR16", "CHR17", "CHR18",
    "CHR19", "CHR20", "CHR21", "CHR22", "CHRX",  "CHRY",  "CHRM"};
// hg19 contig lengths: the concatenated axis the synthetic sources and the shards live on
static const int64_t kContigLength[25] = {
    249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663, 146364022, 141213431,
    135534747, 135006516, 133851895, 115169878, 107349540, 102531392, 90354753,  81195210,  78077248,
    59128983,  63025520,  48129895,  51304566,  155270560, 59373566,  16571};

inline std::string upper(std::string s) {
  for (auto& c : s) c = (char)std::toupper((unsigned char)c);
  return s;
}

struct ParseException : std::runtime_error { using std::runtime_error::runtime_error; };
struct IllegalArgument : std::runtime_error { using std::runtime_error::runtime_error; };
struct CmdLineException : std::runtime_error { using std::runtime_error::runtime_error; };

// Chromosome.valueOf(name.toUpperCase()) — needs the CHR prefix (M/ReadCaller.java:86-87)
inline int chromosome_value_of(const std::string& name) {
  const std::string u = upper(name);
  for (int i = 0; i < 25; i++)
    if (u == kChromosomeNames[i]) return i;
  throw IllegalArgument("No enum constant Chromosome." + u);
}
// Chromosome.fromString — also accepts 1..22, X, Y, M (M/DenovoUtil.java:114-129)
inline int chromosome_from_string(const std::string& str) {
  const std::string u = upper(str);
  for (int i = 0; i < 25; i++)
    if (u == kChromosomeNames[i]) return i;
  for (int i = 0; i < 25; i++)
    if (u == std::string(kChromosomeNames[i]).substr(3)) return i;
  throw IllegalArgument("Unknown Chromosome " + u);
}
inline int64_t contig_offset(int chrom) {
  int64_t off = 0;
  for (int i = 0; i < chrom; i++) off += kContigLength[i];
  return off;
}

// ---- CommandLine (args4j @Option names and defaults, M/CommandLine.java:31-105) ------------
struct CommandLine {
  bool has_caller = false;
  Caller caller = Caller::READ;
  int inferMethod = DNV_METHOD_BAYES;
  std::string outputFileName, outputDir, inputFileName;
  std::string dadCallsetName, momCallsetName, childCallsetName;
  std::string datasetId = "3049512673186936334";
  bool has_dataset = false;
  double sequenceErrorRate = 1e-2, denovoMutationRate = 1e-8, lrtThreshold = 1.0;
  int numThreads = 1;
  LogLevel logLevel = LogLevel::INFO;
  std::string logFile;
  std::vector<std::string> chromosomes;
  bool has_start = false, has_end = false;
  int64_t startPosition = 0, endPosition = 0;
  int64_t maxVariantResults = 1000;
  // synthetic sources (replace the Google Genomics client; not in the reference)
  uint64_t syntheticSeed = 0x5EED0001ULL;
  double syntheticDepth = 30.0;
  std::string syntheticCallset;  // binary dnv_vcall stream, see VariantFileSource
  std::string syntheticSource = "reads";  // reads: synthetic reads -> pileup kernel -> inference; counts: pileup counts directly
  int syntheticReadLength = 100;
  std::vector<int32_t> devices = {0};  // CUDA devices of the READ caller (interval-sharded, records gathered through the first)
  bool help = false;

  static std::string usage() {
    return "Usage: DenovoMain [flags...]\n"
           " --caller <variant|read|full>      : The caller mode\n"
           " --inference_method <map|bayes|lrt>: Inference method (map | bayes | lrt)\n"
           " --output_file <file>              : File to write results\n"
           " --output_dir <dir>                : File to write results\n"
           " --input_calls_file <file>         : File to read from\n"
           " --dad_callset_name <name>         : Dad's callset name e.g. NA12877\n"
           " --mom_callset_name <name>         : Mom's callset name e.g. NA12878\n"
           " --child_callset_name <name>       : Child's callset name e.g. NA12879\n"
           " --dataset_id <id>                 : Dataset id\n"
           " --seq_err_rate <rate>             : Specify the sequence error rate (default 1e-2)\n"
           " --denovo_mut_rate <rate>          : Specify the denovo mutation rate (default 1e-8)\n"
           " --lrt_threshold <sig_level>       : likelihood ratio test significance level (default 1. ;higher the stricter)\n"
           " --num_threads <num>               : Specify the number of threads (default 1 ; 1 to 50 suggested)\n"
           " --log_level <level>               : specify the logging level\n"
           " --log_file <file>                 : specify the log file\n"
           " --chromosome <name>               : specify the chromosomes to search (specify multiple times for multiple chromsomes)\n"
           " --start_position <position>       : start position ( usually 1 )\n"
           " --end_position <position>         : end position ( usually set automatically )\n"
           " --max_variant_results <num>       : max variants returned per request (default 1000)\n"
           "synthetic sources (replace the Google Genomics API):\n"
           " --synthetic_seed <u64>            : seed of the synthetic trio (default 0x5EED0001)\n"
           " --synthetic_depth <x>             : mean read depth per individual (default 30)\n"
           " --synthetic_callset <file>        : binary (variant, call) stream for the VARIANT caller\n"
           " --synthetic_source <reads|counts> : reads (default): aligned reads at the candidates -> pileup kernel -> inference;\n"
           "                                     counts: pileup counts straight from the generator\n"
           " --synthetic_read_length <n>       : length of the synthetic reads (default 100)\n"
           " --device <n>                      : CUDA device ordinal (default 0)\n"
           " --devices <list>                  : CUDA devices, e.g. 0-7 or 0,2,3: candidate sites are sharded by interval over\n"
           "                                     them, the call records gathered through the first\n";
  }

  // "0-7", "0,2,3", "1"
  static std::vector<int32_t> parseDevices(const std::string& v) {
    std::vector<int32_t> out;
    size_t a = 0;
    while (a <= v.size()) {
      size_t b = v.find(',', a);
      if (b == std::string::npos) b = v.size();
      const std::string part = v.substr(a, b - a);
      const size_t dash = part.find('-');
      char* end = nullptr;
      if (part.empty()) throw CmdLineException("\"" + v + "\" is not a valid value for \"--devices\"");
      if (dash == std::string::npos) {
        const long x = std::strtol(part.c_str(), &end, 10);
        if (*end || x < 0) throw CmdLineException("\"" + v + "\" is not a valid value for \"--devices\"");
        out.push_back((int32_t)x);
      } else {
        const long lo = std::strtol(part.substr(0, dash).c_str(), &end, 10);
        const bool bad_lo = *end != 0 || dash == 0;
        const std::string hs = part.substr(dash + 1);
        const long hi = std::strtol(hs.c_str(), &end, 10);
        if (bad_lo || *end || hs.empty() || lo < 0 || hi < lo) throw CmdLineException("\"" + v + "\" is not a valid value for \"--devices\"");
        for (long x = lo; x <= hi; x++) out.push_back((int32_t)x);
      }
      a = b + 1;
    }
    if (out.empty() || out.size() > 64) throw CmdLineException("\"--devices\" takes 1 to 64 devices");
    return out;
  }

  void setArgs(int argc, char** argv) {
    auto need = [&](int& i) -> std::string {
      if (i + 1 >= argc) throw CmdLineException(std::string("Option \"") + argv[i] + "\" takes an operand");
      return argv[++i];
    };
    auto num = [&](const std::string& flag, const std::string& v) -> double {
      char* end = nullptr;
      const double d = std::strtod(v.c_str(), &end);
      if (end == v.c_str() || *end) throw CmdLineException("\"" + v + "\" is not a valid value for \"" + flag + "\"");
      return d;
    };
    auto integer = [&](const std::string& flag, const std::string& v) -> int64_t {
      char* end = nullptr;
      const long long d = std::strtoll(v.c_str(), &end, 0);
      if (end == v.c_str() || *end) throw CmdLineException("\"" + v + "\" is not a valid value for \"" + flag + "\"");
      return d;
    };
    bool has_out = false, has_dad = false, has_mom = false;
    for (int i = 1; i < argc; i++) {
      const std::string f = argv[i];
      if (f == "--help" || f == "-h") { help = true; continue; }
      if (f == "--caller") {
        const std::string v = upper(need(i));  // enum flags parse case-insensitively (README uses `--caller full`)
        if (v == "VARIANT") caller = Caller::VARIANT;
        else if (v == "READ") caller = Caller::READ;
        else if (v == "FULL") caller = Caller::FULL;
        else throw CmdLineException("\"" + v + "\" is not a valid value for \"--caller\"");
        has_caller = true;
      } else if (f == "--inference_method") {
        const std::string v = upper(need(i));
        if (v == "MAP") inferMethod = DNV_METHOD_MAP;
        else if (v == "BAYES") inferMethod = DNV_METHOD_BAYES;
        else if (v == "LRT") inferMethod = DNV_METHOD_LRT;
        else throw CmdLineException("\"" + v + "\" is not a valid value for \"--inference_method\"");
      } else if (f == "--output_file") { outputFileName = need(i); has_out = true; }
      else if (f == "--output_dir") outputDir = need(i);
      else if (f == "--input_calls_file") inputFileName = need(i);
      else if (f == "--dad_callset_name") { dadCallsetName = need(i); has_dad = true; }
      else if (f == "--mom_callset_name") { momCallsetName = need(i); has_mom = true; }
      else if (f == "--child_callset_name") childCallsetName = need(i);
      else if (f == "--dataset_id") { datasetId = need(i); has_dataset = true; }
      else if (f == "--seq_err_rate") sequenceErrorRate = num(f, need(i));
      else if (f == "--denovo_mut_rate") denovoMutationRate = num(f, need(i));
      else if (f == "--lrt_threshold") lrtThreshold = num(f, need(i));
      else if (f == "--num_threads") numThreads = (int)integer(f, need(i));
      else if (f == "--log_level") {
        const std::string v = upper(need(i));
        if (v == "ERROR") logLevel = LogLevel::ERROR;
        else if (v == "INFO") logLevel = LogLevel::INFO;
        else if (v == "DEBUG") logLevel = LogLevel::DEBUG;
        else throw CmdLineException("\"" + v + "\" is not a valid value for \"--log_level\"");
      } else if (f == "--log_file") logFile = need(i);
      else if (f == "--chromosome") chromosomes.push_back(need(i));
      else if (f == "--start_position") { startPosition = integer(f, need(i)); has_start = true; }
      else if (f == "--end_position") { endPosition = integer(f, need(i)); has_end = true; }
      else if (f == "--max_variant_results") maxVariantResults = integer(f, need(i));
      else if (f == "--synthetic_seed") syntheticSeed = (uint64_t)integer(f, need(i));
      else if (f == "--synthetic_depth") syntheticDepth = num(f, need(i));
      else if (f == "--synthetic_callset") syntheticCallset = need(i);
      else if (f == "--synthetic_source") {
        syntheticSource = need(i);
        for (auto& c : syntheticSource) c = (char)std::tolower((unsigned char)c);
        if (syntheticSource != "reads" && syntheticSource != "counts")
          throw CmdLineException("\"" + syntheticSource + "\" is not a valid value for \"--synthetic_source\"");
      }
      else if (f == "--synthetic_read_length") {
        syntheticReadLength = (int)integer(f, need(i));
        if (syntheticReadLength < 1 || syntheticReadLength > 1000) throw CmdLineException("\"--synthetic_read_length\" must lie in 1..1000");
      }
      else if (f == "--device") devices = {(int32_t)integer(f, need(i))};
      else if (f == "--devices") devices = parseDevices(need(i));
      else throw CmdLineException("\"" + f + "\" is not a valid option");
    }
    if (help) return;
    // required = true options (M/CommandLine.java:31-33,39-41,51-57,63-65); the child
    // callset name is NOT required by the reference either.
    if (!has_caller) throw CmdLineException("Option \"--caller\" is required");
    if (!has_out) throw CmdLineException("Option \"--output_file\" is required");
    if (!has_dad) throw CmdLineException("Option \"--dad_callset_name\" is required");
    if (!has_mom) throw CmdLineException("Option \"--mom_callset_name\" is required");
    if (!has_dataset) throw CmdLineException("Option \"--dataset_id\" is required");
  }
};

// ---- logger: "<ERROR|INFO|DEBUG> : msg\n" to console (+ optional file), M/DenovoRunner.java:125-158
class Logger {
 public:
  Logger(LogLevel level, const std::string& file) : level_(level) {
    if (!file.empty()) file_.open(file, std::ios::out | std::ios::trunc);
  }
  void severe(const std::string& m) { log(LogLevel::ERROR, "ERROR", m); }
  void info(const std::string& m) { log(LogLevel::INFO, "INFO", m); }
  void fine(const std::string& m) { log(LogLevel::DEBUG, "DEBUG", m); }
  bool is_fine() const { return level_ >= LogLevel::DEBUG; }

 private:
  void log(LogLevel l, const char* name, const std::string& m) {
    if ((int)l > (int)level_) return;
    const std::string line = std::string(name) + " : " + m + "\n";
    std::fputs(line.c_str(), stderr);
    if (file_.is_open()) { file_ << line; file_.flush(); }
  }
  LogLevel level_;
  std::ofstream file_;
};

constexpr int READ_LOG_FREQ = 100;          // M/DenovoUtil.java:34
constexpr int VARIANT_LOG_FREQ = 1000000;   // M/DenovoUtil.java:35

// ---- synthetic sources --------------------------------------------------------------------
// Poisson CDF thresholds: plain IEEE-double recurrence in a fixed order (identical to
// denovo_b200.engine.poisson_thresholds and tests/oracle_lib.poisson_thresholds).
inline std::vector<uint64_t> poisson_thresholds(double depth) {
  const int kmax = (int)(depth + 12.0 * std::sqrt(depth) + 20.0);
  std::vector<uint64_t> thr;
  double term = std::exp(-depth), cdf = 0.0;
  for (int k = 0; k < kmax; k++) {
    cdf += term;
    const double x = cdf * 18446744073709551616.0;
    thr.push_back(x >= 18446744073709551616.0 ? ~0ULL : (uint64_t)x);
    term = term * depth / (k + 1);
  }
  return thr;
}

// Binary (variant, call) stream for the VARIANT caller.  File layout:
//   char magic[8] = "DNVCALL1"; int32 n_contigs; then per contig:
//   char name[32]; int64 upper_bound; int64 n_records; dnv_vcall records[n_records] (sorted by start)
struct ContigCalls {
  std::string name;
  int64_t upper_bound = 0;
  std::vector<dnv_vcall> records;
};
inline std::vector<ContigCalls> load_callset(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open synthetic callset " + path);
  char magic[8];
  int32_t nc = 0;
  f.read(magic, 8);
  f.read((char*)&nc, 4);
  if (!f || std::memcmp(magic, "DNVCALL1", 8) != 0 || nc < 0) throw std::runtime_error("bad callset file " + path);
  std::vector<ContigCalls> out((size_t)nc);
  for (auto& c : out) {
    char name[32];
    int64_t n = 0;
    f.read(name, 32);
    f.read((char*)&c.upper_bound, 8);
    f.read((char*)&n, 8);
    if (!f || n < 0) throw std::runtime_error("bad callset file " + path);
    name[31] = 0;
    c.name = name;
    c.records.resize((size_t)n);
    f.read((char*)c.records.data(), (std::streamsize)(n * (int64_t)sizeof(dnv_vcall)));
    if (!f) throw std::runtime_error("truncated callset file " + path);
  }
  return out;
}

// ---- DenovoShared (M/DenovoShared.java) ---------------------------------------------------
struct DenovoShared {
  CommandLine cmd;
  std::set<int> chromosomes;  // EnumSet<Chromosome>, ALL when no --chromosome (M/DenovoRunner.java:85-94)
  std::shared_ptr<Logger> logger;
  dnv_ctx* ctx = nullptr;      // single-device engine (VARIANT filter, pileup)
  dnv_group* grp = nullptr;    // the READ caller's devices

  explicit DenovoShared(const CommandLine& c) : cmd(c) {
    if (c.chromosomes.empty()) for (int i = 0; i < 25; i++) chromosomes.insert(i);
    else for (const auto& s : c.chromosomes) chromosomes.insert(chromosome_from_string(s));
    logger = std::make_shared<Logger>(c.logLevel, c.logFile);
  }
  ~DenovoShared() { if (grp) dnv_group_destroy(grp); if (ctx) dnv_destroy(ctx); }
  DenovoShared(const DenovoShared&) = delete;

  // new BayesInfer(shared) -> dnv_create (tables built once, M/ReadCaller.java:49-52)
  dnv_ctx* engine() {
    if (!ctx) {
      dnv_params p{};
      p.seq_err_rate = cmd.sequenceErrorRate;
      p.denovo_mut_rate = cmd.denovoMutationRate;
      p.lrt_threshold = cmd.lrtThreshold;
      p.method = cmd.inferMethod;
      p.device = cmd.devices[0];
      if (dnv_create(&p, &ctx) != DNV_OK) throw std::runtime_error(std::string("dnv_create: ") + dnv_last_error());
    }
    return ctx;
  }
  dnv_params params() const {
    dnv_params p{};
    p.seq_err_rate = cmd.sequenceErrorRate;
    p.denovo_mut_rate = cmd.denovoMutationRate;
    p.lrt_threshold = cmd.lrtThreshold;
    p.method = cmd.inferMethod;
    p.device = cmd.devices[0];
    return p;
  }
  // the reference's one shared BayesInfer, used by every task of the pool (M/ReadCaller.java:49-52,91-95): here one
  // device group — a context per device, shards by interval, records gathered through the first device
  dnv_group* group() {
    if (!grp) {
      const dnv_params p = params();
      if (dnv_group_create(&p, cmd.devices.data(), (int32_t)cmd.devices.size(), &grp) != DNV_OK)
        throw std::runtime_error(std::string("dnv_group_create: ") + dnv_last_error());
    }
    return grp;
  }
};

// page-locked host buffer (dnv_alloc_pinned): what crosses the ABI in bulk lives in pinned memory, so that the
// host-to-device copies run at the PCIe rate and the devices of a group copy concurrently
template <typename T>
struct Pinned {
  T* p = nullptr;
  size_t n = 0;
  explicit Pinned(size_t count) : n(count) {
    void* v = nullptr;
    if (dnv_alloc_pinned(std::max<size_t>(count, 1) * sizeof(T), &v) != DNV_OK)
      throw std::runtime_error(std::string("dnv_alloc_pinned: ") + dnv_last_error());
    p = (T*)v;
  }
  ~Pinned() { dnv_free_pinned(p); }
  Pinned(const Pinned&) = delete;
  Pinned& operator=(const Pinned&) = delete;
  T* data() { return p; }
  T& operator[](size_t i) { return p[i]; }
};

// ---- synthetic read source (bit-for-bit specification: the test oracle's
// synth_ref_reads_at) — stands in for the three reads().search() calls per candidate (M/ReadCaller.java:155-186).
// For every candidate site (global position i) and person, the reads of the synthetic pileup become aligned reads
// of length L that cover the site; everything else in a read is the synthetic reference genome at its own
// position, with '-' gaps.  Output in the layout dnv_pileup takes: sorted by (person, start).
struct SyntheticReads {
  std::vector<int64_t> read_start, base_off;
  std::vector<char> bases;
  std::vector<uint8_t> person;
};
namespace synth {
constexpr uint64_t PHI = 0x9E3779B97F4A7C15ULL;
inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = x + PHI;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline uint64_t key(uint64_t seed, uint64_t stream) { return splitmix64(seed ^ (stream * PHI)); }
}  // namespace synth

inline SyntheticReads synthetic_reads_at(uint64_t seed, const std::vector<int64_t>& index, const std::vector<uint64_t>& thr,
                                         uint32_t err_thr, int L, uint32_t gap_thr) {
  using namespace synth;
  struct R { int64_t start, site; int32_t j; uint8_t person, base; };
  std::vector<R> rd;
  const uint64_t k_class = key(seed, 0), k_ref = key(seed, 1), k_alt = key(seed, 2), k_par = key(seed, 3), k_fill = key(seed, 16);
  for (const int64_t site : index) {
    const uint64_t i = (uint64_t)site;
    const uint32_t c100 = (uint32_t)(((splitmix64(k_class ^ i) >> 32) * 100ULL) >> 32);
    const int cls = c100 < 90 ? 0 : (c100 < 99 ? 1 : 2);
    const int ref = (int)(splitmix64(k_ref ^ i) & 3);
    const int alt = (ref + 1 + (int)(splitmix64(k_alt ^ i) % 3)) & 3;
    const uint64_t pb = splitmix64(k_par ^ i);
    int g[3][2] = {{ref, ref}, {ref, ref}, {ref, ref}};
    if (cls == 1) { g[(int)(pb & 1)][1] = alt; if ((pb >> 1) & 1) g[2][1] = alt; }
    else if (cls == 2) g[2][1] = alt;
    for (int p = 0; p < 3; p++) {
      const uint64_t ud = splitmix64(key(seed, 4 + (uint64_t)p) ^ i);
      const int depth = (int)(std::upper_bound(thr.begin(), thr.end(), ud) - thr.begin());  // #{k : thr[k] <= ud}
      const uint64_t k_read = key(seed, 8 + (uint64_t)p), k_off = key(seed, 12 + (uint64_t)p);
      for (int j = 0; j < depth; j++) {
        const uint64_t rid = (i << 16) | (uint64_t)j;
        const uint64_t u = splitmix64(k_read ^ rid);
        int b = g[p][u & 1];
        if ((uint32_t)(u >> 32) < err_thr) b = (b + 1 + (int)(((u >> 8) & 0xFFFF) % 3)) & 3;
        const int64_t off = (int64_t)(splitmix64(k_off ^ rid) % (uint64_t)L);
        rd.push_back(R{site - off, site, j, (uint8_t)p, (uint8_t)b});
      }
    }
  }
  std::sort(rd.begin(), rd.end(), [](const R& a, const R& b) {
    if (a.person != b.person) return a.person < b.person;
    if (a.start != b.start) return a.start < b.start;
    if (a.site != b.site) return a.site < b.site;
    return a.j < b.j;
  });
  SyntheticReads out;
  const size_t n = rd.size();
  out.read_start.resize(n); out.base_off.resize(n + 1); out.person.resize(n); out.bases.resize(n * (size_t)L);
  const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  auto fill = [&](size_t lo, size_t hi) {
    for (size_t r = lo; r < hi; r++) {
      out.read_start[r] = rd[r].start; out.person[r] = rd[r].person; out.base_off[r] = (int64_t)(r * (size_t)L);
      const uint64_t rid = ((uint64_t)rd[r].site << 16) | (uint64_t)rd[r].j;
      const uint64_t k_gap = key(seed, 17 + 32 * (uint64_t)rd[r].person);
      char* dst = out.bases.data() + r * (size_t)L;
      for (int t = 0; t < L; t++) {
        const int64_t q = rd[r].start + t;
        if (q == rd[r].site) dst[t] = "ACGT"[rd[r].base];
        else if ((uint32_t)(splitmix64(k_gap ^ (rid * 1024 + (uint64_t)t)) >> 32) < gap_thr) dst[t] = '-';
        else dst[t] = "ACGT"[splitmix64(k_fill ^ (uint64_t)q) & 3];
      }
    }
  };
  std::vector<std::thread> th;
  for (unsigned w = 0; w < nt; w++) th.emplace_back(fill, n * w / nt, n * (w + 1) / nt);
  for (auto& t : th) t.join();
  out.base_off[n] = (int64_t)(n * (size_t)L);
  return out;
}

// ---- DenovoCaller (M/DenovoCaller.java:24-78) ----------------------------------------------
class DenovoCaller {
 public:
  virtual ~DenovoCaller() {}
  virtual void execute() = 0;

  struct CallHolder { std::string chromosome; int64_t position; };

  // line.split(",") with Java semantics (trailing empty strings dropped), fields 0 and 1
  static CallHolder parseLine(const std::string& line) {
    std::vector<std::string> parts;
    size_t a = 0;
    for (;;) {
      const size_t b = line.find(',', a);
      parts.push_back(line.substr(a, b == std::string::npos ? std::string::npos : b - a));
      if (b == std::string::npos) break;
      a = b + 1;
    }
    while (!parts.empty() && parts.back().empty()) parts.pop_back();
    if (parts.size() < 2) throw ParseException("Could not parse line : " + line);
    const std::string& p = parts[1];
    char* end = nullptr;
    const long long v = std::strtoll(p.c_str(), &end, 10);
    if (p.empty() || *end || std::isspace((unsigned char)p[0]))
      throw IllegalArgument("For input string: \"" + p + "\"");  // Long.valueOf -> NumberFormatException
    return CallHolder{parts[0], (int64_t)v};
  }
  static void writeCalls(std::ofstream& w, const std::string& calls) { w << calls; w.flush(); }
};

// readCounts=CHILD:{...};MOM:{...};DAD:{...},maxGenoType=[D, M, C],isDenovo=b  (M/BayesInfer.java:57-68)
// counts as TreeMap.toString with only alleles seen at least once (M/ReadSummary.java:27,55-57)
inline std::string format_read_details(const uint16_t counts[12], int dad, int mom, int child, bool is_denovo) {
  auto person = [&](int p) {  // packed order 0 DAD, 1 MOM, 2 CHILD
    std::string s = "{";
    bool first = true;
    for (int a = 0; a < 4; a++)
      if (counts[p * 4 + a]) {
        if (!first) s += ", ";
        s += kAlleleNames[a];
        s += "=" + std::to_string(counts[p * 4 + a]);
        first = false;
      }
    return s + "}";
  };
  return "readCounts=CHILD:" + person(2) + ";MOM:" + person(1) + ";DAD:" + person(0) + ",maxGenoType=[" +
         kGenotypeNames[dad] + ", " + kGenotypeNames[mom] + ", " + kGenotypeNames[child] +
         "],isDenovo=" + (is_denovo ? "true" : "false");
}
// PositionCall.toString: [pos,{CHILD=g, MOM=g, DAD=g}]  (M/VariantsBuffer.java:358-360)
inline std::string format_position_call(const dnv_pcall& c) {
  return "[" + std::to_string(c.position) + ",{CHILD=" + kGenotypeNames[c.child] + ", MOM=" +
         kGenotypeNames[c.mom] + ", DAD=" + kGenotypeNames[c.dad] + "}]";
}

// ---- GpuReadCaller: ReadCaller.execute (M/ReadCaller.java:57-103) as ONE batched call ------
class GpuReadCaller : public DenovoCaller {
 public:
  explicit GpuReadCaller(DenovoShared& shared) : shared_(shared) {}
  void execute() override {
    auto& log = *shared_.logger;
    log.info("---- Starting Bayesian Read Caller ----");
    std::ifstream in(shared_.cmd.inputFileName);
    if (!in) throw std::runtime_error(shared_.cmd.inputFileName + " (No such file or directory)");
    std::ofstream out(shared_.cmd.outputFileName, std::ios::out | std::ios::trunc);
    if (!out) throw std::runtime_error(shared_.cmd.outputFileName + " (cannot open for writing)");
    std::vector<CallHolder> sites;
    std::vector<int64_t> index;
    std::string line;
    int lineCount = 0;
    while (std::getline(in, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      const CallHolder h = parseLine(line);
      if (lineCount % READ_LOG_FREQ == 0 && lineCount > 0)
        log.info(std::to_string(lineCount) + " Read candidates processed");
      lineCount++;
      const int chrom = chromosome_value_of(h.chromosome);  // throws like Chromosome.valueOf
      if (!shared_.chromosomes.count(chrom)) continue;      // :86-89
      sites.push_back(h);
      index.push_back(contig_offset(chrom) + h.position);
    }
    const int64_t n = (int64_t)sites.size();
    if (n > 0) {
      dnv_ctx* ctx = shared_.engine();
      const auto thr = poisson_thresholds(shared_.cmd.syntheticDepth);
      const uint32_t err_thr = (uint32_t)(shared_.cmd.sequenceErrorRate * 4294967296.0);
      Pinned<uint16_t> counts((size_t)n * 12);
      if (shared_.cmd.syntheticSource == "counts") {
        // pileup counts straight from the generator (no reads)
        if (dnv_synth_counts_at(ctx, shared_.cmd.syntheticSeed, index.data(), n, thr.data(), (int32_t)thr.size(), err_thr,
                                counts.data()) != DNV_OK)
          throw std::runtime_error(std::string("synthetic pileup source: ") + dnv_last_error());
      } else {
        // getReadMap + getReadSummaryMap (M/ReadCaller.java:113-116,138-186): the reads at every candidate, then
        // `new ReadSummary(reads, position)` x3 for ALL candidates in one dnv_pileup call.  The pileup wants its
        // positions ascending: candidates are looked up through their sorted, de-duplicated positions.
        std::vector<int64_t> uniq(index);
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        const SyntheticReads reads = synthetic_reads_at(shared_.cmd.syntheticSeed, uniq, thr, err_thr,
                                                        shared_.cmd.syntheticReadLength, (uint32_t)(1e-3 * 4294967296.0));
        std::vector<uint16_t> ucounts(uniq.size() * 12);
        if (dnv_pileup(ctx, reads.read_start.data(), reads.base_off.data(), reads.bases.data(), reads.person.data(),
                       (int64_t)reads.read_start.size(), uniq.data(), (int64_t)uniq.size(), ucounts.data()) != DNV_OK)
          throw std::runtime_error(std::string("dnv_pileup: ") + dnv_last_error());
        for (int64_t i = 0; i < n; i++) {
          const size_t u = (size_t)(std::lower_bound(uniq.begin(), uniq.end(), index[(size_t)i]) - uniq.begin());
          std::memcpy(counts.data() + (size_t)i * 12, ucounts.data() + u * 12, 12 * sizeof(uint16_t));
        }
      }
      // BayesInfer.infer for every candidate (M/ReadCaller.java:119-120): ONE call, interval-sharded over --devices
      Pinned<dnv_call> calls((size_t)n);
      int64_t n_calls = 0;
      if (dnv_group_infer_counts(shared_.group(), counts.data(), n, calls.data(), n, &n_calls) != DNV_OK)
        throw std::runtime_error(std::string("dnv_group_infer_counts: ") + dnv_last_error());
      std::string buf;
      for (int64_t i = 0; i < n_calls; i++) {
        const dnv_call& c = calls[(size_t)i];
        const CallHolder& h = sites[(size_t)c.site];
        const std::string l = h.chromosome + "," + std::to_string(h.position) + "," +
                              format_read_details(c.counts, c.dad, c.mom, c.child, true);
        if (log.is_fine()) log.fine(l);
        buf += l + "\n";  // "%s,%d,%s%n", :126-127
      }
      writeCalls(out, buf);
    }
    log.info("---- Read caller terminated ----");
  }

 private:
  DenovoShared& shared_;
};

// ---- GpuVariantCaller: VariantCaller.execute / callSimpleDenovo (M/VariantCaller.java:55-210)
class GpuVariantCaller : public DenovoCaller {
 public:
  explicit GpuVariantCaller(DenovoShared& shared) : shared_(shared) {}
  void execute() override {
    auto& log = *shared_.logger;
    log.info("---- Starting Variant Caller ----");
    if (shared_.cmd.syntheticCallset.empty())
      throw std::runtime_error("the VARIANT caller needs --synthetic_callset <file> (the Google Genomics API is not available)");
    const auto contigs = load_callset(shared_.cmd.syntheticCallset);
    std::ofstream out(shared_.cmd.outputFileName, std::ios::out | std::ios::trunc);
    if (!out) throw std::runtime_error(shared_.cmd.outputFileName + " (cannot open for writing)");
    int64_t variantCounter = 0;
    for (const auto& contig : contigs) {
      if (!shared_.chromosomes.count(chromosome_value_of(contig.name))) continue;  // :75-84
      const int64_t startContigPos = shared_.cmd.has_start ? shared_.cmd.startPosition : 1;
      const int64_t endContigPos = shared_.cmd.has_end ? shared_.cmd.endPosition : contig.upper_bound;
      const int nt = std::max(1, shared_.cmd.numThreads);
      const int64_t stride = (endContigPos - startContigPos) / nt;  // :88-90
      log.info("Processing Chromosome : " + contig.name);
      for (int t = 0; t < nt; t++) {  // one fresh VariantsBuffer per position shard, :92-102,125
        const int64_t start = startContigPos + t * stride;
        const int64_t end = t == nt - 1 ? endContigPos : start + stride - 1;
        callSimpleDenovo(out, contig, start, end, variantCounter);
      }
    }
    log.info("---- Variant caller terminated ----");
  }

 private:
  // The variant source returns every variant overlapping [start, end] (the API's overlap query,
  // M/VariantContigStream.java:40-47): start <= reqEnd and end > reqStart.
  void callSimpleDenovo(std::ofstream& out, const ContigCalls& contig, int64_t start, int64_t end,
                        int64_t& variantCounter) {
    std::vector<dnv_vcall> shard;
    int32_t last_vid = -1;
    bool first = true;
    for (const auto& r : contig.records)
      if (r.start <= end && r.end > start) {
        shard.push_back(r);
        if (first || r.variant_id != last_vid) {
          if (variantCounter % VARIANT_LOG_FREQ == 0 && variantCounter > 0)
            shared_.logger->info(std::to_string(variantCounter) + " Variant candidates processed");
          variantCounter++;
          last_vid = r.variant_id;
          first = false;
        }
      }
    if (shard.empty()) return;
    std::vector<dnv_pcall> pc(shard.size());
    int64_t n_out = 0;
    if (dnv_mendel_filter(shared_.engine(), shard.data(), (int64_t)shard.size(), pc.data(), (int64_t)pc.size(), &n_out) != DNV_OK)
      throw std::runtime_error(std::string("dnv_mendel_filter: ") + dnv_last_error());
    std::string buf;
    for (int64_t i = 0; i < n_out; i++)
      if (pc[(size_t)i].is_denovo) {  // :176-183
        const std::string l = contig.name + "," + std::to_string(pc[(size_t)i].position) + "," + format_position_call(pc[(size_t)i]);
        if (shared_.logger->is_fine()) shared_.logger->fine(l);
        buf += l + "\n";
      }
    writeCalls(out, buf);
  }
  DenovoShared& shared_;
};

// ---- DenovoCallers factory (M/DenovoCallers.java:32-45) -------------------------------------
struct DenovoCallers {
  static std::unique_ptr<DenovoCaller> getVariantCaller(DenovoShared& s) { return std::unique_ptr<DenovoCaller>(new GpuVariantCaller(s)); }
  static std::unique_ptr<DenovoCaller> getReadCaller(DenovoShared& s) { return std::unique_ptr<DenovoCaller>(new GpuReadCaller(s)); }
};

// ---- DenovoRunner.execute (M/DenovoRunner.java:165-186) -------------------------------------
inline void run(const CommandLine& cmd) {
  if (cmd.caller == Caller::VARIANT) {
    DenovoShared shared(cmd);
    DenovoCallers::getVariantCaller(shared)->execute();
  } else if (cmd.caller == Caller::READ) {
    if (cmd.inputFileName.empty()) throw IllegalArgument("Input calls file needed for read mode");  // M/DenovoRunner.java:171-172
    DenovoShared shared(cmd);
    DenovoCallers::getReadCaller(shared)->execute();
  } else {  // FULL: VARIANT -> <out>.tmp, then READ over it (:173-182)
    CommandLine stage1 = cmd;
    stage1.caller = Caller::VARIANT;
    stage1.outputFileName = cmd.outputFileName + ".tmp";
    run(stage1);
    CommandLine stage2 = cmd;
    stage2.caller = Caller::READ;
    stage2.inputFileName = cmd.outputFileName + ".tmp";
    run(stage2);
  }
}

}  // namespace denovo
