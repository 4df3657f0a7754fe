// strugatzki_host.hpp -- C++17 host side above the C ABI (include/strugatzki_b200.h).
//
// The reference is compiled JVM code (Scala) and no JDK exists in the build image, so this header is the
// compiled-language mirror of the reference's operator interface for the hot path: the same Config fields,
// defaults and XML tags (Api/FeatureCorrelation.scala:105-273, Api/FeatureSegmentation.scala:71-191,
// Api/SelfSimilarity.scala:61-283, Api/FeatureExtraction.scala:163-206), the same Processor contract
// (start / abort / progress / observer receiving Progress and Result(Success | Failure(Aborted) | Failure(e)),
// Strugatzki.scala:95-99,177-211), and processor bodies that do what `...Impl.body()` does up to the arithmetic
// (read meta XML, list the database folder, read feat_norms.aif and the feature AIFFs) and then call libsgz_b200.so.
// Header only; link with -lsgz_b200.  There is no CPU fallback: without a B200 every processor fails with the
// library's error message.
#pragma once

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <mutex>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "strugatzki_b200.h"

namespace strugatzki {

struct Aborted : std::runtime_error {            // Processor.Aborted
  Aborted() : std::runtime_error("Aborted") {}
};

// ---------------------------------------------------------------------------------------------------------
// Span (de.sciss.span.Span.NonVoid): start / stop in sample frames, either may be open
// ---------------------------------------------------------------------------------------------------------
struct Span {
  std::optional<int64_t> start, stop;
  Span() = default;
  Span(int64_t a, int64_t b) : start(a), stop(b) {}
  static Span all() { return Span(); }
  static Span from(int64_t a) { Span s; s.start = a; return s; }
  static Span until(int64_t b) { Span s; s.stop = b; return s; }
  bool operator==(const Span &o) const { return start == o.start && stop == o.stop; }
};
inline int64_t spacing(const Span &a, const Span &b) {     // SpanUtil.spacing, SpanUtil.scala:38-43
  return *a.start < *b.start ? *b.start - *a.stop : *a.start - *b.stop;
}

// ---------------------------------------------------------------------------------------------------------
// minimal XML helpers (flat elements with unique tags, one nesting level for <punchIn>/<punchOut>/<span>)
// ---------------------------------------------------------------------------------------------------------
namespace xml {
inline std::string escape(const std::string &s) {
  std::string o;
  for (char c : s) { if (c == '&') o += "&amp;"; else if (c == '<') o += "&lt;"; else if (c == '>') o += "&gt;"; else o += c; }
  return o;
}
inline std::string unescape(std::string s) {
  auto rep = [&](const char *a, const char *b) { size_t p = 0; while ((p = s.find(a, p)) != std::string::npos) { s.replace(p, strlen(a), b); p += strlen(b); } };
  rep("&lt;", "<"); rep("&gt;", ">"); rep("&amp;", "&");
  return s;
}
inline std::optional<std::string> child(const std::string &x, const std::string &tag) {
  const std::string open = "<" + tag + ">", close = "</" + tag + ">", empty = "<" + tag + "/>";
  size_t a = x.find(open);
  if (a == std::string::npos) { if (x.find(empty) != std::string::npos) return std::string(); return std::nullopt; }
  size_t b = x.find(close, a);
  if (b == std::string::npos) throw std::runtime_error("malformed XML: <" + tag + ">");
  return unescape(x.substr(a + open.size(), b - a - open.size()));
}
inline std::string text(const std::string &x, const std::string &tag) {
  auto c = child(x, tag);
  if (!c) throw std::runtime_error("missing XML element <" + tag + ">");
  std::string s = *c;
  size_t i = s.find_first_not_of(" \t\r\n"), j = s.find_last_not_of(" \t\r\n");
  return i == std::string::npos ? std::string() : s.substr(i, j - i + 1);
}
inline std::string el(const std::string &tag, const std::string &v) { return "<" + tag + ">" + escape(v) + "</" + tag + ">"; }
template <typename T> std::string el(const std::string &tag, T v) { std::ostringstream o; o.precision(9); o << v; return el(tag, o.str()); }
inline std::string el(const std::string &tag, bool v) { return el(tag, std::string(v ? "true" : "false")); }
inline std::string span_children(const Span &s) {
  std::string o;
  if (s.start) o += el("start", *s.start);
  if (s.stop) o += el("stop", *s.stop);
  return o;
}
inline Span span_from(const std::optional<std::string> &x) {
  Span s;
  if (!x) return s;
  if (auto a = child(*x, "start")) s.start = std::stoll(*a);
  if (auto b = child(*x, "stop")) s.stop = std::stoll(*b);
  return s;
}
inline std::string slurp(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("In file: " + path);
  std::ostringstream o; o << f.rdbuf(); return o.str();
}
}  // namespace xml

inline int fullToFeat(int64_t n, int step) { return (int)((n + (step >> 1)) / step); }   // FeatureCorrelationImpl.scala:38

// ---------------------------------------------------------------------------------------------------------
// FeatureExtraction.Config meta file (persisted fields only; the extractor is out of scope)
// ---------------------------------------------------------------------------------------------------------
struct FeatureExtractionConfig {
  std::string audioInput = "input.aif", featureOutput = "features.aif";
  std::optional<std::string> metaOutput;
  int numCoeffs = 13, fftSize = 1024, fftOverlap = 2, channelsBehavior = 0;
  int stepSize() const { return fftSize / fftOverlap; }
  std::string toXML() const {
    return "<feature>" + xml::el("input", audioInput) + xml::el("output", featureOutput) +
           xml::el("meta", metaOutput.value_or("")) + xml::el("numCoeffs", numCoeffs) + xml::el("fftSize", fftSize) +
           xml::el("fftOverlap", fftOverlap) + xml::el("channels", channelsBehavior) + "</feature>";
  }
  static FeatureExtractionConfig fromXML(const std::string &x) {
    FeatureExtractionConfig c;
    c.audioInput = xml::text(x, "input");
    c.featureOutput = xml::text(x, "output");
    std::string m = xml::child(x, "meta").value_or("");
    if (!m.empty()) c.metaOutput = m;
    c.numCoeffs = std::stoi(xml::text(x, "numCoeffs"));
    c.fftSize = std::stoi(xml::text(x, "fftSize"));
    c.fftOverlap = std::stoi(xml::text(x, "fftOverlap"));
    std::string ch = xml::child(x, "channels").value_or("");
    c.channelsBehavior = ch.empty() ? 0 : std::stoi(ch);
    if (c.channelsBehavior < 0 || c.channelsBehavior > 2) throw std::invalid_argument(ch);
    return c;
  }
  static FeatureExtractionConfig fromXMLFile(const std::string &path) { return fromXML(xml::slurp(path)); }
  bool operator==(const FeatureExtractionConfig &o) const {
    return audioInput == o.audioInput && featureOutput == o.featureOutput && metaOutput == o.metaOutput &&
           numCoeffs == o.numCoeffs && fftSize == o.fftSize && fftOverlap == o.fftOverlap && channelsBehavior == o.channelsBehavior;
  }
};

// ---------------------------------------------------------------------------------------------------------
// AIFF-C 'fl32' feature files: the raw big-endian payload is handed to the GPU (SGZ_LAYOUT_INTERLEAVED_BE)
// ---------------------------------------------------------------------------------------------------------
struct FeatureFile {
  int numChannels = 0;
  int64_t numFrames = 0;
  std::vector<unsigned char> payload;   // big-endian float32, interleaved
};
inline uint32_t be32(const unsigned char *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline FeatureFile readFeatureFile(const std::string &path) {
  std::string b = xml::slurp(path);
  const unsigned char *u = (const unsigned char *)b.data();
  if (b.size() < 12 || memcmp(u, "FORM", 4) != 0 || (memcmp(u + 8, "AIFC", 4) != 0 && memcmp(u + 8, "AIFF", 4) != 0))
    throw std::runtime_error(path + ": not an AIFF file");
  FeatureFile f;
  size_t pos = 12, ssndOff = 0, ssndLen = 0;
  bool fl32 = false;
  while (pos + 8 <= b.size()) {
    uint32_t sz = be32(u + pos + 4);
    if (memcmp(u + pos, "COMM", 4) == 0) {
      f.numChannels = (u[pos + 8] << 8) | u[pos + 9];
      f.numFrames = be32(u + pos + 10);
      fl32 = sz >= 22 && (memcmp(u + pos + 26, "fl32", 4) == 0 || memcmp(u + pos + 26, "FL32", 4) == 0);
    } else if (memcmp(u + pos, "SSND", 4) == 0) {
      ssndOff = pos + 16 + be32(u + pos + 8);
      ssndLen = sz - 8;
    }
    pos += 8 + sz + (sz & 1);
  }
  if (f.numFrames > 0 && !fl32) throw std::runtime_error(path + ": feature files must be float32 AIFF-C (fl32)");
  size_t need = (size_t)f.numFrames * f.numChannels * 4;
  if (need > ssndLen || ssndOff + need > b.size()) throw std::runtime_error(path + ": truncated SSND chunk");
  f.payload.assign(u + ssndOff, u + ssndOff + need);
  return f;
}
// feat_norms.aif -> [numCh][2] = {min, max}; `require` of FeatureCorrelationImpl.scala:61-71
inline std::vector<float> readNormFile(const std::string &databaseFolder, int numCh) {
  FeatureFile f = readFeatureFile((std::filesystem::path(databaseFolder) / "feat_norms.aif").string());
  if (!(f.numChannels == numCh && f.numFrames == 2)) throw std::invalid_argument("requirement failed: feat_norms.aif shape");
  std::vector<float> n((size_t)numCh * 2);
  for (int fr = 0; fr < 2; fr++)
    for (int c = 0; c < numCh; c++) {
      uint32_t v = be32(f.payload.data() + 4 * ((size_t)fr * numCh + c));
      memcpy(&n[2 * c + fr], &v, 4);
    }
  return n;
}

// ---------------------------------------------------------------------------------------------------------
// Processor contract
// ---------------------------------------------------------------------------------------------------------
struct Event {
  enum Kind { Progress, Success, Failure, AbortedResult } kind;
  double progress = 0;
  std::string message;
};

template <typename Product>
class ProcessorImpl {
 public:
  virtual ~ProcessorImpl() { if (thread_.joinable()) thread_.join(); }
  void addListener(std::function<void(const Event &)> fn) { listeners_.push_back(std::move(fn)); }
  void start() {
    if (thread_.joinable()) throw std::logic_error("processor already started");
    thread_ = std::thread([this] {
      Event ev;
      try { product_ = body(); ev.kind = Event::Success; }
      catch (const Aborted &) { ev.kind = Event::AbortedResult; failed_ = std::make_exception_ptr(Aborted()); }
      catch (const std::exception &e) { ev.kind = Event::Failure; ev.message = e.what(); failed_ = std::current_exception(); }
      done_ = true;
      for (auto &l : listeners_) l(ev);
    });
  }
  void abort() { aborted_ = true; onAbort(); }
  bool isCompleted() const { return done_; }
  double progress() const { return progress_; }
  Product await() {                          // Await.result(processor, Duration.Inf)
    if (thread_.joinable()) thread_.join();
    if (failed_) std::rethrow_exception(failed_);
    return product_;
  }

 protected:
  virtual Product body() = 0;
  virtual void onAbort() {}
  void checkAborted() const { if (aborted_) throw Aborted(); }
  void setProgress(double p) {
    progress_ = p;
    if (p - lastDispatched_ >= 0.01 || (p >= 1.0 && lastDispatched_ < 1.0)) {
      lastDispatched_ = p;
      Event ev; ev.kind = Event::Progress; ev.progress = p;
      for (auto &l : listeners_) l(ev);
    }
  }
  static void check(int rc) {
    if (rc >= 0) return;
    if (rc == SGZ_ERR_ABORTED) throw Aborted();
    if (rc == SGZ_ERR_INVALID) throw std::invalid_argument(sgz_last_error());
    throw std::runtime_error(sgz_last_error());
  }
  std::atomic<bool> aborted_{false}, done_{false};

 private:
  std::vector<std::function<void(const Event &)>> listeners_;
  std::thread thread_;
  std::atomic<double> progress_{0.0};
  double lastDispatched_ = -1.0;
  Product product_{};
  std::exception_ptr failed_;
};

// ---------------------------------------------------------------------------------------------------------
// FeatureCorrelation
// ---------------------------------------------------------------------------------------------------------
namespace FeatureCorrelation {

struct Punch {                                   // Api/FeatureCorrelation.scala:93-100
  Span span{0, 44100};
  float temporalWeight = 0.5f;
  std::string toXML(const std::string &tag) const {
    return "<" + tag + ">" + xml::span_children(span) + xml::el("weight", temporalWeight) + "</" + tag + ">";
  }
  static Punch fromXML(const std::string &x) { Punch p; p.span = xml::span_from(x); p.temporalWeight = std::stof(xml::text(x, "weight")); return p; }
  bool operator==(const Punch &o) const { return span == o.span && temporalWeight == o.temporalWeight; }
};

struct Match {                                   // Api/FeatureCorrelation.scala:54-70
  float sim = 0;
  std::string file;
  Span punch;
  float boostIn = 1, boostOut = 1;
  std::string toXML() const {
    return "<match>" + xml::el("sim", sim) + xml::el("file", file) + xml::el("start", *punch.start) + xml::el("stop", *punch.stop) +
           xml::el("boostIn", boostIn) + xml::el("boostOut", boostOut) + "</match>";
  }
  static Match fromXML(const std::string &x) {
    Match m;
    m.sim = std::stof(xml::text(x, "sim")); m.file = xml::text(x, "file");
    m.punch = Span(std::stoll(xml::text(x, "start")), std::stoll(xml::text(x, "stop")));
    m.boostIn = std::stof(xml::text(x, "boostIn")); m.boostOut = std::stof(xml::text(x, "boostOut"));
    return m;
  }
  bool operator==(const Match &o) const { return sim == o.sim && file == o.file && punch == o.punch && boostIn == o.boostIn && boostOut == o.boostOut; }
};

struct Config {                                  // ConfigBuilder defaults, :168-206
  std::string databaseFolder = "database", metaInput = "input_feat.xml";
  Punch punchIn;
  std::optional<Punch> punchOut;
  int64_t minPunch = 22050, maxPunch = 88200;
  bool normalize = true;
  float maxBoost = 8.f;
  int numMatches = 1, numPerFile = 1;
  int64_t minSpacing = 0;
  std::string toXML() const {
    return "<correlate>" + xml::el("database", databaseFolder) + xml::el("input", metaInput) + punchIn.toXML("punchIn") +
           (punchOut ? punchOut->toXML("punchOut") : std::string()) + xml::el("minPunch", minPunch) + xml::el("maxPunch", maxPunch) +
           xml::el("normalize", normalize) + xml::el("maxBoost", maxBoost) + xml::el("numMatches", numMatches) +
           xml::el("numPerFile", numPerFile) + xml::el("minSpacing", minSpacing) + "</correlate>";
  }
  static Config fromXML(const std::string &x) {
    Config c;
    c.databaseFolder = xml::text(x, "database"); c.metaInput = xml::text(x, "input");
    c.punchIn = Punch::fromXML(*xml::child(x, "punchIn"));
    if (auto po = xml::child(x, "punchOut")) c.punchOut = Punch::fromXML(*po);
    c.minPunch = std::stoll(xml::text(x, "minPunch")); c.maxPunch = std::stoll(xml::text(x, "maxPunch"));
    c.normalize = xml::text(x, "normalize") == "true"; c.maxBoost = std::stof(xml::text(x, "maxBoost"));
    c.numMatches = std::stoi(xml::text(x, "numMatches")); c.numPerFile = std::stoi(xml::text(x, "numPerFile"));
    c.minSpacing = std::stoll(xml::text(x, "minSpacing"));
    return c;
  }
  bool operator==(const Config &o) const {
    return databaseFolder == o.databaseFolder && metaInput == o.metaInput && punchIn == o.punchIn && punchOut == o.punchOut &&
           minPunch == o.minPunch && maxPunch == o.maxPunch && normalize == o.normalize && maxBoost == o.maxBoost &&
           numMatches == o.numMatches && numPerFile == o.numPerFile && minSpacing == o.minSpacing;
  }
};

// DB discovery of FeatureCorrelationImpl.scala:42-55; order = sorted file name (the reference iterates a HashSet)
inline std::vector<FeatureExtractionConfig> listDatabase(const Config &c, const FeatureExtractionConfig &in) {
  namespace fs = std::filesystem;
  std::vector<std::string> names;
  for (auto &e : fs::directory_iterator(c.databaseFolder)) {
    std::string n = e.path().filename().string();
    if (n.size() >= 9 && n.compare(n.size() - 9, 9, "_feat.xml") == 0 &&
        fs::absolute(e.path()).lexically_normal() != fs::absolute(c.metaInput).lexically_normal())
      names.push_back(e.path().string());
  }
  std::sort(names.begin(), names.end());
  std::vector<FeatureExtractionConfig> out;
  for (auto &n : names) {
    FeatureExtractionConfig e = FeatureExtractionConfig::fromXMLFile(n);
    if (e.numCoeffs == in.numCoeffs && e.fftSize / e.fftOverlap == in.stepSize()) out.push_back(e);
  }
  return out;
}

class Processor : public ProcessorImpl<std::vector<Match>> {
 public:
  explicit Processor(Config c, int device = 0) : config(std::move(c)), device_(device) {}
  const Config config;

 protected:
  void onAbort() override { if (sgz_corr *j = job_.load()) sgz_corr_abort(j); }
  std::vector<Match> body() override {
    const FeatureExtractionConfig extrIn = FeatureExtractionConfig::fromXMLFile(config.metaInput);
    const int step = extrIn.stepSize(), numCh = extrIn.numCoeffs + 1;
    const std::vector<FeatureExtractionConfig> dbs = listDatabase(config, extrIn);
    std::vector<float> norm;
    if (config.normalize) norm = readNormFile(config.databaseFolder, numCh);
    FeatureFile in = readFeatureFile(extrIn.featureOutput);
    checkAborted();
    sgz_ctx *ctx = nullptr; sgz_db *db = nullptr; sgz_corr *job = nullptr;
    struct Guard { sgz_ctx *&c; sgz_db *&d; sgz_corr *&j; std::atomic<sgz_corr *> &slot;
                   ~Guard() { slot = nullptr; if (j) sgz_corr_destroy(j); if (d) sgz_db_destroy(d); if (c) sgz_ctx_destroy(c); } } guard{ctx, db, job, job_};
    check(sgz_ctx_create(device_, &ctx));
    check(sgz_db_create(ctx, numCh, config.normalize ? norm.data() : nullptr, &db));
    for (const auto &e : dbs) {
      checkAborted();
      FeatureFile f = readFeatureFile(e.featureOutput);
      if (f.numChannels != numCh) throw std::runtime_error(e.featureOutput + ": channel count mismatch");
      check(sgz_db_add_file(db, f.payload.data(), f.numFrames, SGZ_LAYOUT_INTERLEAVED_BE));
    }
    check(sgz_db_finalize(db));
    sgz_corr_config cc{};
    cc.stepSize = step;
    cc.punchInStart = *config.punchIn.span.start; cc.punchInStop = *config.punchIn.span.stop;
    cc.punchInWeight = config.punchIn.temporalWeight;
    cc.hasPunchOut = config.punchOut ? 1 : 0;
    if (config.punchOut) { cc.punchOutStart = *config.punchOut->span.start; cc.punchOutStop = *config.punchOut->span.stop; cc.punchOutWeight = config.punchOut->temporalWeight; }
    cc.minPunch = config.minPunch; cc.maxPunch = config.maxPunch; cc.maxBoost = config.maxBoost;
    cc.numMatches = config.numMatches; cc.numPerFile = config.numPerFile; cc.minSpacing = config.minSpacing;
    check(sgz_corr_create(db, &cc, in.payload.data(), in.numFrames, SGZ_LAYOUT_INTERLEAVED_BE, &job));
    job_ = job;
    check(sgz_corr_start(job));
    for (;;) {
      float p = 0; int32_t done = 0, status = 0;
      check(sgz_corr_poll(job, &p, &done, &status));
      setProgress(std::min<double>(p, 0.999));
      if (done) { check(status); break; }
      if (aborted_) sgz_corr_abort(job);
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    check(sgz_corr_wait(job));
    int32_t n = 0;
    check(sgz_corr_result(job, nullptr, 0, &n));
    std::vector<sgz_match> raw((size_t)std::max(n, 1));
    check(sgz_corr_result(job, raw.data(), (int32_t)raw.size(), &n));
    std::vector<Match> out;
    for (int i = 0; i < n; i++) {
      Match m;
      m.sim = raw[i].sim; m.file = dbs[raw[i].file].audioInput; m.punch = Span(raw[i].start, raw[i].stop);
      m.boostIn = raw[i].boostIn; m.boostOut = raw[i].boostOut;
      out.push_back(m);
    }
    setProgress(1.0);
    return out;
  }

 private:
  int device_;
  std::atomic<sgz_corr *> job_{nullptr};
};
}  // namespace FeatureCorrelation

// ---------------------------------------------------------------------------------------------------------
// FeatureSegmentation
// ---------------------------------------------------------------------------------------------------------
namespace FeatureSegmentation {
struct Break {
  float sim = 0; int64_t pos = 0;
  std::string toXML() const { return "<break>" + xml::el("sim", sim) + xml::el("pos", pos) + "</break>"; }
  static Break fromXML(const std::string &x) { Break b; b.sim = std::stof(xml::text(x, "sim")); b.pos = std::stoll(xml::text(x, "pos")); return b; }
  bool operator==(const Break &o) const { return sim == o.sim && pos == o.pos; }
};
struct Config {                                  // Api/FeatureSegmentation.scala:134-159
  std::string databaseFolder = "database", metaInput = "input_feat.xml";
  Span span = Span::all();
  int64_t corrLen = 22050;
  float temporalWeight = 0.5f;
  bool normalize = true;
  int numBreaks = 1;
  int64_t minSpacing = 22050;
  std::string toXML() const {
    return "<segmentation>" + xml::el("database", databaseFolder) + xml::el("input", metaInput) + "<span>" + xml::span_children(span) +
           "</span>" + xml::el("corr", corrLen) + xml::el("weight", temporalWeight) + xml::el("normalize", normalize) +
           xml::el("numBreaks", numBreaks) + xml::el("minSpacing", minSpacing) + "</segmentation>";
  }
  static Config fromXML(const std::string &x) {
    Config c;
    c.databaseFolder = xml::text(x, "database"); c.metaInput = xml::text(x, "input"); c.span = xml::span_from(xml::child(x, "span"));
    c.corrLen = std::stoll(xml::text(x, "corr")); c.temporalWeight = std::stof(xml::text(x, "weight"));
    c.normalize = xml::text(x, "normalize") == "true"; c.numBreaks = std::stoi(xml::text(x, "numBreaks"));
    c.minSpacing = std::stoll(xml::text(x, "minSpacing"));
    return c;
  }
  bool operator==(const Config &o) const {
    return databaseFolder == o.databaseFolder && metaInput == o.metaInput && span == o.span && corrLen == o.corrLen &&
           temporalWeight == o.temporalWeight && normalize == o.normalize && numBreaks == o.numBreaks && minSpacing == o.minSpacing;
  }
};
class Processor : public ProcessorImpl<std::vector<Break>> {
 public:
  explicit Processor(Config c, int device = 0) : config(std::move(c)), device_(device) {}
  const Config config;

 protected:
  std::vector<Break> body() override {
    const FeatureExtractionConfig extr = FeatureExtractionConfig::fromXMLFile(config.metaInput);
    const int numCh = extr.numCoeffs + 1;
    std::vector<float> norm;
    if (config.normalize) norm = readNormFile(config.databaseFolder, numCh);
    FeatureFile f = readFeatureFile(extr.featureOutput);
    checkAborted();
    sgz_ctx *ctx = nullptr;
    check(sgz_ctx_create(device_, &ctx));
    sgz_segm_config sc{};
    sc.stepSize = extr.stepSize(); sc.hasStart = config.span.start ? 1 : 0; sc.hasStop = config.span.stop ? 1 : 0;
    sc.spanStart = config.span.start.value_or(0); sc.spanStop = config.span.stop.value_or(0); sc.corrLen = config.corrLen;
    sc.temporalWeight = config.temporalWeight; sc.numBreaks = config.numBreaks; sc.minSpacing = config.minSpacing;
    std::vector<sgz_break> raw((size_t)std::max(config.numBreaks, 0) + 1);
    int32_t n = 0; int64_t nOff = 0;
    int rc = sgz_segm_run(ctx, &sc, numCh, config.normalize ? norm.data() : nullptr, f.payload.data(), f.numFrames,
                          SGZ_LAYOUT_INTERLEAVED_BE, raw.data(), (int32_t)raw.size(), &n, nullptr, 0, &nOff);
    sgz_ctx_destroy(ctx);
    check(rc);
    std::vector<Break> out;
    for (int i = 0; i < n; i++) { Break b; b.sim = raw[i].sim; b.pos = raw[i].pos; out.push_back(b); }
    setProgress(1.0);
    return out;
  }

 private:
  int device_;
};
}  // namespace FeatureSegmentation

// ---------------------------------------------------------------------------------------------------------
// SelfSimilarity (PNG written with stored-deflate blocks: no zlib dependency)
// ---------------------------------------------------------------------------------------------------------
namespace SelfSimilarity {
enum class ColorScheme { GrayScale, PsychoOptical };
struct Config {                                  // Api/SelfSimilarity.scala:153-187
  std::string databaseFolder = "database", metaInput = "input_feat.xml";
  std::optional<std::string> metaInput2;
  std::string imageOutput = "output_selfsim.png";
  Span span = Span::all();
  int64_t corrLen = 44100;
  int decimation = 1;
  float temporalWeight = 0.5f;
  ColorScheme colors = ColorScheme::PsychoOptical;
  float colorWarp = 1.f, colorCeil = 1.f;
  bool colorInv = false, normalize = true;
  std::string toXML() const {
    return "<selfsimilarity>" + xml::el("database", databaseFolder) + xml::el("input", metaInput) +
           (metaInput2 ? xml::el("input2", *metaInput2) : std::string()) + xml::el("output", imageOutput) +
           ((span.start || span.stop) ? "<span>" + xml::span_children(span) + "</span>" : std::string()) + xml::el("corr", corrLen) +
           xml::el("decimation", decimation) + xml::el("weight", temporalWeight) +
           xml::el("colors", std::string(colors == ColorScheme::GrayScale ? "gray" : "psycho")) + xml::el("colorWarp", colorWarp) +
           xml::el("colorCeil", colorCeil) + xml::el("colorInv", colorInv) + xml::el("normalize", normalize) + "</selfsimilarity>";
  }
  static Config fromXML(const std::string &x) {
    Config c;
    c.databaseFolder = xml::text(x, "database"); c.metaInput = xml::text(x, "input");
    if (auto i2 = xml::child(x, "input2")) c.metaInput2 = *i2;
    c.imageOutput = xml::text(x, "output"); c.span = xml::span_from(xml::child(x, "span"));
    c.corrLen = std::stoll(xml::text(x, "corr")); c.decimation = std::stoi(xml::text(x, "decimation"));
    c.temporalWeight = std::stof(xml::text(x, "weight"));
    std::string col = xml::text(x, "colors");
    if (col == "gray") c.colors = ColorScheme::GrayScale; else if (col == "psycho") c.colors = ColorScheme::PsychoOptical;
    else throw std::invalid_argument("MatchError: " + col);
    c.colorWarp = std::stof(xml::text(x, "colorWarp")); c.colorCeil = std::stof(xml::text(x, "colorCeil"));
    c.colorInv = xml::text(x, "colorInv") == "true"; c.normalize = xml::text(x, "normalize") == "true";
    return c;
  }
  bool operator==(const Config &o) const {
    return databaseFolder == o.databaseFolder && metaInput == o.metaInput && metaInput2 == o.metaInput2 && imageOutput == o.imageOutput &&
           span == o.span && corrLen == o.corrLen && decimation == o.decimation && temporalWeight == o.temporalWeight &&
           colors == o.colors && colorWarp == o.colorWarp && colorCeil == o.colorCeil && colorInv == o.colorInv && normalize == o.normalize;
  }
};

inline void writePNG(const std::string &path, const int32_t *rgb, int w, int h) {
  static uint32_t crcTab[256];
  static bool init = false;
  if (!init) { for (uint32_t n = 0; n < 256; n++) { uint32_t c = n; for (int k = 0; k < 8; k++) c = c & 1 ? 0xedb88320u ^ (c >> 1) : c >> 1; crcTab[n] = c; } init = true; }
  auto crc = [&](const std::string &d) { uint32_t c = 0xffffffffu; for (unsigned char ch : d) c = crcTab[(c ^ ch) & 0xff] ^ (c >> 8); return c ^ 0xffffffffu; };
  auto be = [](uint32_t v) { std::string s(4, '\0'); s[0] = (char)(v >> 24); s[1] = (char)(v >> 16); s[2] = (char)(v >> 8); s[3] = (char)v; return s; };
  auto chunk = [&](const std::string &tag, const std::string &data) { return be((uint32_t)data.size()) + tag + data + be(crc(tag + data)); };
  std::string raw;
  raw.reserve((size_t)h * (3 * (size_t)w + 1));
  for (int y = 0; y < h; y++) {
    raw.push_back('\0');
    for (int x = 0; x < w; x++) { int32_t p = rgb[(size_t)y * w + x]; raw.push_back((char)(p >> 16)); raw.push_back((char)(p >> 8)); raw.push_back((char)p); }
  }
  std::string z = "\x78\x01";                     // zlib header, stored blocks
  uint32_t a = 1, b2 = 0;
  for (unsigned char ch : raw) { a = (a + ch) % 65521u; b2 = (b2 + a) % 65521u; }
  for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
    size_t n = std::min<size_t>(65535, raw.size() - off);
    bool last = off + n >= raw.size();
    z.push_back(last ? 1 : 0);
    z.push_back((char)(n & 0xff)); z.push_back((char)(n >> 8)); z.push_back((char)(~n & 0xff)); z.push_back((char)((~n >> 8) & 0xff));
    z.append(raw, off, n);
    if (last) break;
  }
  z += be((b2 << 16) | a);
  std::string ihdr = be((uint32_t)w) + be((uint32_t)h) + std::string("\x08\x02\x00\x00\x00", 5);
  std::ofstream f(path, std::ios::binary);
  f << std::string("\x89PNG\r\n\x1a\n", 8) << chunk("IHDR", ihdr) << chunk("IDAT", z) << chunk("IEND", "");
}

class Processor : public ProcessorImpl<int> {
 public:
  explicit Processor(Config c, int device = 0, std::vector<int32_t> palette = {}, bool precise = false)
      : config(std::move(c)), device_(device), palette_(std::move(palette)), precise_(precise) {}
  const Config config;

 protected:
  int body() override {
    const FeatureExtractionConfig e1 = FeatureExtractionConfig::fromXMLFile(config.metaInput);
    const FeatureExtractionConfig e2 = config.metaInput2 ? FeatureExtractionConfig::fromXMLFile(*config.metaInput2) : e1;
    if (!(e1.fftSize == e2.fftSize && e1.fftOverlap == e2.fftOverlap && e1.numCoeffs == e2.numCoeffs)) throw std::invalid_argument("requirement failed");
    if (config.colors == ColorScheme::PsychoOptical && palette_.empty())
      throw std::runtime_error("PsychoOptical needs de.sciss.intensitypalette's table (pass it as `palette`) or use GrayScale");
    const int numCh = e1.numCoeffs + 1;
    std::vector<float> norm;
    if (config.normalize) norm = readNormFile(config.databaseFolder, numCh);
    FeatureFile f1 = readFeatureFile(e1.featureOutput), f2;
    const bool cross = e1.featureOutput != e2.featureOutput;
    if (cross) f2 = readFeatureFile(e2.featureOutput);
    checkAborted();
    sgz_self_config sc{};
    sc.stepSize = e1.stepSize(); sc.hasStart = config.span.start ? 1 : 0; sc.hasStop = config.span.stop ? 1 : 0;
    sc.spanStart = config.span.start.value_or(0); sc.spanStop = config.span.stop.value_or(0); sc.corrLen = config.corrLen;
    sc.decimation = config.decimation; sc.temporalWeight = config.temporalWeight; sc.colorInv = config.colorInv;
    sc.colorWarp = config.colorWarp; sc.colorCeil = config.colorCeil;
    sc.lut = config.colors == ColorScheme::PsychoOptical ? palette_.data() : nullptr;
    sc.lutSize = config.colors == ColorScheme::PsychoOptical ? (int32_t)palette_.size() : 0;
    sc.precise = precise_ ? 1 : 0;
    sgz_self_geometry g{};
    check(sgz_self_geometry_of(&sc, f1.numFrames, cross ? f2.numFrames : f1.numFrames, &g));
    std::vector<int32_t> rgb((size_t)std::max(g.imgExt, 1) * std::max(g.imgExt, 1));
    sgz_ctx *ctx = nullptr;
    check(sgz_ctx_create(device_, &ctx));
    int rc = sgz_self_run(ctx, &sc, numCh, config.normalize ? norm.data() : nullptr, f1.payload.data(), f1.numFrames,
                          cross ? f2.payload.data() : nullptr, cross ? f2.numFrames : 0, SGZ_LAYOUT_INTERLEAVED_BE, 0, 0,
                          rgb.data(), (int64_t)rgb.size(), &g);
    sgz_ctx_destroy(ctx);
    check(rc);
    checkAborted();
    writePNG(config.imageOutput, rgb.data(), g.imgExt, g.imgExt);       // ImageIO.write, SelfSimilarityImpl.scala:167
    setProgress(1.0);
    return g.imgExt;
  }

 private:
  int device_;
  std::vector<int32_t> palette_;
  bool precise_;
};
}  // namespace SelfSimilarity

}  // namespace strugatzki
