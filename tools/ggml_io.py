"""Reader for the legacy ggml `.bin` Whisper format - test/tool side only."""
import struct

import numpy as np


def read_ggml(path):
    with open(path, "rb") as f:
        magic, = struct.unpack("<I", f.read(4))
        assert magic == 0x67676D6C
        hp = struct.unpack("<11i", f.read(44))
        keys = ("n_vocab", "n_audio_ctx", "n_audio_state", "n_audio_head", "n_audio_layer",
                "n_text_ctx", "n_text_state", "n_text_head", "n_text_layer", "n_mels", "ftype")
        hparams = dict(zip(keys, hp))
        n_mel, n_fft = struct.unpack("<2i", f.read(8))
        filters = np.frombuffer(f.read(4 * n_mel * n_fft), np.float32).reshape(n_mel, n_fft).copy()
        n_vocab, = struct.unpack("<i", f.read(4))
        vocab = []
        for _ in range(n_vocab):
            ln, = struct.unpack("<I", f.read(4))
            vocab.append(f.read(ln))
        tensors = {}
        while True:
            hd = f.read(12)
            if len(hd) < 12:
                break
            n_dims, name_len, ttype = struct.unpack("<3i", hd)
            ne = struct.unpack("<%di" % n_dims, f.read(4 * n_dims))
            name = f.read(name_len).decode()
            shape = tuple(reversed(ne))
            n = int(np.prod(shape))
            if ttype == 0:
                a = np.frombuffer(f.read(4 * n), np.float32).reshape(shape).copy()
            elif ttype == 1:
                a = np.frombuffer(f.read(2 * n), np.float16).reshape(shape).astype(np.float32)
            elif ttype in (2, 3, 6, 7, 8):  # q4_0, q4_1, q5_0, q5_1, q8_0: dequantised by the gguf package
                from gguf import GGMLQuantizationType, quants
                bb = {2: 18, 3: 20, 6: 22, 7: 24, 8: 34}[ttype]
                raw = np.frombuffer(f.read(n // 32 * bb), np.uint8).reshape(shape[:-1] + (shape[-1] // 32 * bb,))
                a = quants.dequantize(raw, GGMLQuantizationType(ttype)).astype(np.float32)
            else:
                raise ValueError("unsupported ggml tensor type %d" % ttype)
            tensors[name] = a
    return hparams, filters, vocab, tensors
