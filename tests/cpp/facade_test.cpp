// Compiles against include/afs_b200.hpp and links libafs_b200.so: a reference-style caller.
// Without a CUDA device it must fail loudly (no CPU fallback); with one it prints formants and audio
// statistics that tests/test_facade.py compares with the ctypes path.
#include <cmath>
#include <cstdio>

#include "afs_b200.hpp"

int main() {
  try {
    afs::Device dev(0);
    // --- TlModel-style use (Data::getVowelFormants, Data.cpp:149-196) ---
    const double a_params[16] = {2.066256, 1.143496, 0.117848, 2.693147, 5.744530, 6.962660, 8.276990, 0.940841,
                                 1.252275, 14.184641, 6.665092, 0.243414, 16.520416, 2.830321, 16.529516, 2.968984};
    afs::TlModel tl(dev);
    tl.tubes.resize(2);
    afs::calculateOneDimTubeFunction(a_params, &tl.tubes[0]);
    tl.tubes[0].setGlottisArea(0.0);
    tl.tubes[1] = tl.tubes[0];
    tl.tubes[1].setVelumOpening(0.5);
    std::vector<double> f, bw;
    std::vector<int32_t> n, flags;
    tl.getFormants(&f, &bw, &n, 8, &flags);
    std::printf("formants %d %.6f %.6f %.6f nasal %d %.6f\n", n[0], f[0], f[1], f[2], flags[1] & 4 ? 1 : 0, f[8]);
    afs::ComplexSpectrum s;
    tl.getSpectrum(afs::TlModel::FLOW_SOURCE_TF, &s, 2048, afs::Tube::FIRST_PHARYNX_SECTION);
    std::printf("spectrum %d %.17g %.17g\n", s.numFreq, s.re[0], s.im[0]);
    std::vector<double> h;
    tl.getImpulseResponse(&h, 9);  // Synthesizer::synthesizeVowelLf (Synthesizer.cpp:719-724)
    std::printf("impulse %zu %.17g %.17g\n", h.size(), h[100], h[512 + 100]);
    // --- Data::optimizeFormantsVowel-style use (Data.cpp:203-487), three runs ---
    std::vector<double> prm(a_params, a_params + 16);
    const std::vector<double> err = afs::optimizeFormantsVowel(tl, &prm, {650.0, 1300.0, 2600.0}, 1.0, 0.0, nullptr, 3);
    std::printf("optimizer %.12g %.12g %.12g\n", err[0], prm[2], prm[8]);
    // --- Synthesizer::synthesizeVowelLf / playTargetSequence-style use ---
    afs_lf_pulse lf;
    afs_lf_pulse_default(&lf);
    const std::vector<int16_t> pcm = afs::synthesizeVowelsLf(tl, {lf, lf}, false);
    long long e16 = 0;
    for (int16_t v : pcm) e16 += (long long)v * v;
    std::printf("lfvowels %zu %lld\n", pcm.size(), e16);
    {
      afs_tds_options to;
      afs_tds_options_default(&to);
      to.generate_noise_sources = 0;
      std::vector<double> shapes(64);
      for (int k = 0; k < 4; k++) for (int j = 0; j < 16; j++) shapes[16 * k + j] = a_params[j] * (j == 7 ? 1.0 + 0.1 * k : 1.0);
      const double st[4] = {0.02, 0.01, 0.01, 0.02}, tr[3] = {0.02, 0.02, 0.02}, g6[6] = {120, 8000, 0.01, 0.01, 0, -40};
      const std::vector<double> a = afs::playTargetSequences(dev, to, shapes, st, tr, g6);
      double et = 0;
      for (double v : a) et += v * v;
      std::printf("targets %zu %.12e\n", a.size(), et);
    }
    // --- Synthesizer-style use (Synthesizer.cpp:515-639): prime, then one 50 ms frame ---
    afs::TdsSynthesizer syn(dev, 2);
    afs::Tube tubes[2] = {tl.tubes[0], tl.tubes[1]};
    const double g[12] = {120, 8000, 0.01, 0.01, 0, -40, 150, 8000, 0.01, 0.01, 0, -40};
    syn.synthesizeSignalTds(tubes, g, 1102);
    syn.synthesizeSignalTds(tubes, g, 1102);
    std::vector<double> audio;
    std::vector<uint32_t> status;
    syn.flush(&audio, nullptr, &status);
    double e = 0;
    for (double v : audio) e += v * v;
    std::printf("audio %zu %.12e status %u\n", audio.size(), e, status[0] | status[1]);
    // --- TdsModel::setFlowSource / getSectionPressure / getSectionFlow, TlModel::getMeanFlow ---
    syn.setFlowSource(50.0, 30);
    syn.setProbeSections({24, 40});
    syn.synthesizeSignalTds(tubes, g, 200);
    syn.synthesizeSignalTds(tubes, g, 200);
    syn.flush(&audio);
    double ep = 0, ei = 0;
    for (double v : syn.sectionPressure()) ep += v * v;
    for (size_t i = 0; i < syn.sectionInflow().size(); i++) ei += syn.sectionInflow()[i] * syn.sectionOutflow()[i];
    std::printf("probes %zu %.12e %.12e\n", syn.sectionPressure().size(), ep, ei);
    const std::vector<double> mf = tl.getMeanFlow(8000.0);
    const std::vector<int32_t> mc = tl.getMostConstrictedSection();
    std::printf("meanflow %.17g %d\n", mf[0], mc[0]);
    // --- the same utterance in two pieces on models that live across calls (afs_tds_stream) ---
    {
      afs::TdsSynthesizer chunked(dev, 2);
      std::vector<double> a1, a2;
      chunked.synthesizeSignalTds(tubes, g, 0);
      chunked.synthesizeSignalTds(tubes, g, 551);
      chunked.flushAndContinue(&a1);
      const std::vector<unsigned char> blob = chunked.checkpoint();
      chunked.synthesizeSignalTds(tubes, g, 551);
      chunked.flushAndContinue(&a2);
      afs::TdsSynthesizer resumed(dev, 2);   // a new object continues from the checkpoint, bit for bit
      std::vector<double> a3;
      resumed.restore(blob);
      resumed.synthesizeSignalTds(tubes, g, 551);
      resumed.flushAndContinue(&a3);
      if (a3 != a2) { std::printf("error 99 restored stream differs\n"); return 2; }
      double ce = 0;
      for (double v : a1) ce += v * v;
      for (double v : a2) ce += v * v;
      std::printf("stream %zu %.12e\n", a1.size() + a2.size(), ce);
    }
    // --- the same utterances over every visible device (afs_group) ---
    afs::DeviceGroup grp;
    afs::TdsSynthesizer gsyn(grp, 2);
    gsyn.synthesizeSignalTds(tubes, g, 1102);
    gsyn.synthesizeSignalTds(tubes, g, 1102);
    std::vector<double> gaudio;
    gsyn.flush(&gaudio);
    double ge = 0;
    for (double v : gaudio) ge += v * v;
    std::printf("group %d %zu %.12e\n", grp.size(), gaudio.size(), ge);
  } catch (const afs::Error &err) {
    std::printf("error %d %s\n", err.status, err.what());
    return err.status == AFS_ERR_NO_DEVICE ? 0 : 2;
  }
  return 0;
}
