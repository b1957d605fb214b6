"""Segment post-filter: host/text_filters.h restates is_hallucination
(/root/reference/src/utils.h:214-306, applied at stt_engine.cpp:272-278). Pinned against the reference's own code
(oracle/_ref/libref_wav.so compiles utils.h where it lies) on a generated corpus: every phrase of the rule
tables in several casings / with punctuation and context, bracketed and punctuation-only segments, fillers,
random ASCII and UTF-8 text."""
import ctypes as C
import os
import random
import subprocess

import pytest

from conftest import PKG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SEEDS = ["altyazı", "Altyazı M.K.", "sesli betimleme", "senkron", "www.example.com", "izlediğiniz için teşekkürler",
         "teşekkür ederim", "thank you", "Thanks for watching", "abone ol", "videoyu beğen", "bir sonraki videoda",
         "devam edecek", "transcription:", "subtitle:", "2分", "ご視聴ありがとうございました", "I'm going to go", "Okay.",
         "Bye.", "Ahem", "Umarım", "Hıhı", "Pffft", "Ehem", "Hmm", "Aa", "Ah", "Oh", "Eh", "[music]", "(applause)",
         "...", " . ", "a", "", "hello world", "Okay, let us start the meeting.", "The senkr report", "oh no",
         "Ohio is a state", "ah!", "Hmm...", "eh?", "visit www", ".com", "thanks", "Thank you very much.",
         "merhaba nasılsınız", "bugün hava çok güzel", "çğıöşü ÇĞİÖŞÜ", "  padded  ", "\tTabbed\n", "[unclosed", "closed)"]


def corpus():
    rng = random.Random(5)
    out = list(SEEDS)
    wraps = ["%s", " %s ", "%s.", "%s!", "%s?", "%s...", "...%s", "(%s)", "[%s]", "so %s then", "%s, right", "\"%s\""]
    for s in SEEDS:
        for w in wraps:
            out.append(w % s)
            out.append((w % s).upper())
            out.append((w % s).lower())
            out.append((w % s).title())
    alphabet = "abcdefghijklmnopqrstuvwxyz ABCDEFGHIJKLM.,!?[]()'\"-0123456789çğıöşüİ分ご"
    for _ in range(1500):
        out.append("".join(rng.choice(alphabet) for _ in range(rng.randint(0, 24))))
    return out


@pytest.fixture(scope="module")
def libs():
    subprocess.check_call(["make", "-s", "-C", PKG])
    subprocess.check_call(["make", "-s", "-C", os.path.join(PKG, "host")])
    ours = C.CDLL(os.path.join(PKG, "libstt_engine.so")).stt_is_hallucination
    ours.argtypes = [C.c_char_p]
    path = os.path.join(ROOT, "oracle", "_ref", "libref_wav.so")
    if not os.path.exists(path) and os.path.exists("/root/reference/src/utils.h"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    ref = None
    if os.path.exists(path):
        ref = C.CDLL(path).ref_is_hallucination
        ref.argtypes = [C.c_char_p]
    return ours, ref


def test_known_decisions(libs):
    ours, _ = libs
    for s in ("", "a", " . ", "[music]", "(applause)", "Thanks for watching", "Altyazı M.K.", "Okay.", "Hmm", "ehem..."):
        assert ours(s.encode()) == 1, s
    for s in ("hello world", "Okay, let us start the meeting.", "The senkr report", "visit www.example.org"):
        assert ours(s.encode()) == 0, s


def test_matches_the_reference_build_on_a_corpus(libs):
    ours, ref = libs
    if ref is None:
        pytest.skip("oracle/_ref/libref_wav.so not built here (no /root/reference)")
    texts = corpus()
    diff = [t for t in texts if b"\0" not in t.encode() and ours(t.encode()) != ref(t.encode())]
    assert len(texts) > 3000 and not diff, diff[:20]
