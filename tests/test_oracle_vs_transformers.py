"""Live cross-check of the oracle against transformers.MimiModel (third-party home of the reference's
arithmetic) where it is importable. Small input so the CPU suite stays fast."""
import numpy as np
import pytest

from oracle import mimi_oracle as O
from tokenize_audio_b200 import synth

transformers = pytest.importorskip("transformers")
torch = pytest.importorskip("torch")


def test_oracle_equals_mimimodel(state_dict):
    from transformers import MimiConfig, MimiModel
    model = MimiModel(MimiConfig()).eval()
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in state_dict.items()}, strict=False)
    assert not unexpected
    x = np.stack([synth.synth_speech(21, 9000), synth.synth_speech(22, 9000)])[:, None, :]
    with torch.no_grad():
        ref = model.encode(torch.from_numpy(x), num_quantizers=32).audio_codes.numpy()
        lat = model.downsample(model.encoder_transformer(model.encoder(torch.from_numpy(x)).transpose(1, 2))[0].transpose(1, 2)).numpy()
    taps = {}
    codes = O.encode(state_dict, x, 32, taps=taps)
    assert (codes == ref).mean() >= 0.999
    assert np.linalg.norm(np.stack(taps["latent"]) - lat) / np.linalg.norm(lat) < 1e-5
    # synthetic codebooks are filled (random-init MimiModel has all-zero embed_sum -> all codes 0)
    assert len(np.unique(ref)) > 50
