/* CPU restatement of the reference's per-segment prosody DSP and speaker clustering.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, bench.py's cpu_baseline leg). The product path never links it.
 *
 * Follows /root/reference/src/prosody_extractor.cpp:31-224 (extract_prosody) and
 * /root/reference/src/speaker_cluster.cpp:5-38 (SpeakerClusterer), called per segment from
 * /root/reference/src/stt_engine.cpp:313-334.
 * PINNED: unlike the Whisper arithmetic this part of the reference is plain C++ in the tree, so
 * oracle/Makefile compiles the reference's own two source files where they lie into
 * oracle/_ref/libref_prosody.so (through ref_prosody_shim.cpp) and tests/test_prosody.py checks this
 * restatement against it bit for bit on seeded inputs and against tests/golden/prosody_ref.npz
 * (generated from the reference build by tests/golden/make_prosody_golden.py).
 * Floating point: IEEE single precision without fused multiply-add contraction (the reference's
 * build flags, -O2/-O3 for generic x86-64, do not contract); this file is compiled with
 * -ffp-contract=off and the CUDA kernels use __fmul_rn/__fadd_rn for the same reason. */
#pragma once
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_prosody_opts {  /* ProsodyOptions, prosody_extractor.h:20-26 */
  float lpf_alpha, gender_threshold, min_pitch, max_pitch;
} ora_prosody_opts;

typedef struct ora_prosody {       /* AffectiveTags, prosody_extractor.h:6-18 */
  char gender;                     /* 'M', 'F', '?' */
  int emotion;                     /* 0 neutral, 1 excited, 2 sad, 3 angry */
  float arousal, valence, pitch_mean, pitch_std, energy_mean, energy_std, spectral_centroid,
      zero_crossing_rate;
  float speaker_vec[8];
} ora_prosody;

void ora_prosody_extract(const float* pcm, size_t n_samples, int sample_rate, const ora_prosody_opts* opts,
                         ora_prosody* out);
/* online clustering of n 8-D vectors in order; ids[i] = index k of "spk_k" */
void ora_speaker_cluster(const float* vecs, int n, float threshold, int* ids);

#ifdef __cplusplus
}
#endif
