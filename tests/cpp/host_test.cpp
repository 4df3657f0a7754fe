// Test driver of the C++ host layer (host/strugatzki_host.hpp).
//   host_test xml                         -> the reference's own test (StrugatzkiSuite.scala:12-92): XML round trips
//   host_test corr  <db> <meta> K npf minSpacing [poStart poStop minPunch maxPunch]
//   host_test segm  <db> <meta> numBreaks
//   host_test self  <db> <meta> <png> decim precise
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "strugatzki_host.hpp"

using namespace strugatzki;

#define REQUIRE(c) do { if (!(c)) { fprintf(stderr, "REQUIRE failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

static int test_xml() {
  FeatureExtractionConfig fe;
  fe.audioInput = "/abs/testing.aif"; fe.featureOutput = "relative.aif"; fe.numCoeffs += 1; fe.fftSize += 1; fe.fftOverlap += 1;
  FeatureExtractionConfig fe2 = fe; fe2.metaOutput = "/abs";
  REQUIRE(FeatureExtractionConfig::fromXML(fe.toXML()) == fe);
  REQUIRE(FeatureExtractionConfig::fromXML(fe2.toXML()) == fe2);

  FeatureCorrelation::Config fc;
  fc.databaseFolder = "/abs/db"; fc.metaInput = "rarara.xml";
  fc.punchIn.span = Span(*fc.punchIn.span.start + 1, *fc.punchIn.span.stop + 2); fc.punchIn.temporalWeight += 0.11f;
  fc.punchOut = FeatureCorrelation::Punch{Span(555, 666), 0.1234f};
  fc.minPunch += 1; fc.maxPunch += 2; fc.normalize = !fc.normalize; fc.maxBoost += 1; fc.numMatches += 1; fc.numPerFile += 1;
  fc.minSpacing += 1;
  FeatureCorrelation::Config fc2 = fc; fc2.punchOut.reset(); fc2.normalize = !fc2.normalize;
  REQUIRE(FeatureCorrelation::Config::fromXML(fc.toXML()) == fc);
  REQUIRE(FeatureCorrelation::Config::fromXML(fc2.toXML()) == fc2);
  REQUIRE(!(fc == fc2));

  FeatureCorrelation::Match m1{0.23f, "gaga.aif", Span(33, 44), -6.f, -7.f}, m2{0.46f, "/abs/rara.wav", Span(666, 777), 8.f, 9.f};
  REQUIRE(FeatureCorrelation::Match::fromXML(m1.toXML()) == m1);
  REQUIRE(FeatureCorrelation::Match::fromXML(m2.toXML()) == m2);

  FeatureSegmentation::Config fs;
  fs.databaseFolder = "/abs/db"; fs.metaInput = "rarara.xml"; fs.span = Span(1, 2); fs.corrLen += 1; fs.temporalWeight += 0.1f;
  fs.normalize = !fs.normalize; fs.numBreaks += 1; fs.minSpacing += 1;
  FeatureSegmentation::Config fs2 = fs; fs2.span = Span::all(); fs2.normalize = !fs2.normalize;
  REQUIRE(FeatureSegmentation::Config::fromXML(fs.toXML()) == fs);
  REQUIRE(FeatureSegmentation::Config::fromXML(fs2.toXML()) == fs2);
  FeatureSegmentation::Break b{0.5f, 12345};
  REQUIRE(FeatureSegmentation::Break::fromXML(b.toXML()) == b);

  SelfSimilarity::Config ss;
  ss.metaInput2 = "other_feat.xml"; ss.span = Span::from(1000); ss.decimation = 3; ss.colors = SelfSimilarity::ColorScheme::GrayScale;
  ss.colorWarp = 0.5f; ss.colorCeil = 0.75f; ss.colorInv = true; ss.normalize = false;
  REQUIRE(SelfSimilarity::Config::fromXML(ss.toXML()) == ss);
  REQUIRE(SelfSimilarity::Config::fromXML(SelfSimilarity::Config().toXML()) == SelfSimilarity::Config());

  // defaults are the reference's
  FeatureCorrelation::Config d;
  REQUIRE(d.databaseFolder == "database" && d.metaInput == "input_feat.xml" && d.punchIn.span == Span(0, 44100) &&
          d.punchIn.temporalWeight == 0.5f && !d.punchOut && d.minPunch == 22050 && d.maxPunch == 88200 && d.normalize &&
          d.maxBoost == 8.f && d.numMatches == 1 && d.numPerFile == 1 && d.minSpacing == 0);
  REQUIRE(spacing(Span(0, 10), Span(15, 20)) == 5 && spacing(Span(0, 10), Span(5, 20)) == -5);
  REQUIRE(fullToFeat(88200, 512) == 172 && fullToFeat(22050, 512) == 43);
  printf("xml ok\n");
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  try {
    if (mode == "xml") return test_xml();
    if (mode == "corr" && argc >= 7) {
      FeatureCorrelation::Config c;
      c.databaseFolder = argv[2]; c.metaInput = argv[3];
      c.punchIn = FeatureCorrelation::Punch{Span(0, 88200), 0.5f};
      c.numMatches = atoi(argv[4]); c.numPerFile = atoi(argv[5]); c.minSpacing = atoll(argv[6]);
      if (argc >= 11) { c.punchOut = FeatureCorrelation::Punch{Span(atoll(argv[7]), atoll(argv[8])), 0.5f}; c.minPunch = atoll(argv[9]); c.maxPunch = atoll(argv[10]); }
      FeatureCorrelation::Processor p(c);
      int progressEvents = 0, results = 0;
      p.addListener([&](const Event &e) { if (e.kind == Event::Progress) progressEvents++; else results++; });
      p.start();
      auto res = p.await();
      for (auto &m : res) printf("match %.9g %s %lld %lld %.9g %.9g\n", m.sim, m.file.c_str(), (long long)*m.punch.start,
                                 (long long)*m.punch.stop, m.boostIn, m.boostOut);
      printf("events %d %d\n", progressEvents, results);
      return 0;
    }
    if (mode == "segm" && argc >= 5) {
      FeatureSegmentation::Config c;
      c.databaseFolder = argv[2]; c.metaInput = argv[3]; c.numBreaks = atoi(argv[4]);
      FeatureSegmentation::Processor p(c);
      p.start();
      for (auto &b : p.await()) printf("break %.9g %lld\n", b.sim, (long long)b.pos);
      return 0;
    }
    if (mode == "self" && argc >= 7) {
      SelfSimilarity::Config c;
      c.databaseFolder = argv[2]; c.metaInput = argv[3]; c.imageOutput = argv[4]; c.decimation = atoi(argv[5]);
      c.colors = SelfSimilarity::ColorScheme::GrayScale; c.corrLen = 20480;
      SelfSimilarity::Processor p(c, 0, {}, atoi(argv[6]) != 0);
      p.start();
      printf("imgExt %d\n", p.await());
      return 0;
    }
  } catch (const Aborted &) {
    printf("aborted\n");
    return 3;
  } catch (const std::exception &e) {
    printf("failure: %s\n", e.what());
    return 4;
  }
  return 2;
}
