"""TEST INFRASTRUCTURE — CPU restatement of the reference's CosyVoice2 streaming vocoder orchestration.  Only tests/ may import it.

fade_in_out: cosyvoice/utils/common.py:169-177.  token2wav_cv2: the vocoder half of CosyVoice2Model.token2wav
(cosyvoice/cli/model.py:290-313) driven over the per-call mel slices `tts_mel[:, :, token_offset * token_mel_ratio:]`:
8 cached mel frames are prepended, the source signal of the previous chunk's last 3840 samples is handed to the vocoder as
`cache_source`, the chunk head is cross-faded with the previous chunk's held-back tail and every non-final chunk holds its own
last 3840 samples back.  Parity unpinned by the reference (it has no tests); pinned to the reference FUNCTION in
tests/test_host_cpu.py when /root/reference is present."""
import numpy as np
import torch


def fade_in_out(fade_in_mel, fade_out_mel, window):
    """window: numpy float64 (np.hamming(2 * overlap)); the blend is evaluated in float64 and stored into the float32 tensor"""
    n = int(window.shape[0] / 2)
    out = fade_in_mel.clone()
    out[..., :n] = out[..., :n] * window[:n] + fade_out_mel[..., -n:] * window[n:]
    return out


def token2wav_cv2(vocoder, mels, mel_cache_len=8, frame_samples=480):
    """vocoder(mel (1, C, T), cache_source (1, 1, n)) -> (speech (1, T*frame), source (1, 1, T*frame)); mels: the new mel frames
    of every token2wav call, the last one being the finalize=True call.  Returns the list of emitted speech chunks."""
    src_len = mel_cache_len * frame_samples
    window = np.hamming(2 * src_len)
    cache, out = None, []
    for i, mel in enumerate(mels):
        finalize = i == len(mels) - 1
        if cache is not None:
            mel = torch.cat([cache["mel"], mel], dim=2)
            src_cache = cache["source"]
        else:
            src_cache = torch.zeros(1, 1, 0)
        speech, source = vocoder(mel, src_cache)
        if cache is not None:
            speech = fade_in_out(speech, cache["speech"], window)
        if not finalize:
            cache = {"mel": mel[:, :, -mel_cache_len:], "source": source[:, :, -src_len:], "speech": speech[:, -src_len:]}
            speech = speech[:, :-src_len]
        out.append(speech)
    return out
