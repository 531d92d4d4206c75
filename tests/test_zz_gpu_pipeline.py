"""GPU: the host-buffer frame pipeline (transformerupscaler_b200.pipeline.FramePipeline: copy-in / copy-out streams and alternating
compute streams, the path bench.py's `e2e` number goes through) returns, batch for batch, exactly the bytes of a plain synchronous
call of the same model -- also when the forwards of consecutive batches overlap on two streams.  (Replaces the synchronous
.to(device) -> model -> .cpu() loops of speed_test.py:60-67 / app_overlay.py:365-391.)"""
import importlib

import pytest
import torch

from oracle.weights import synth_state_dict, synth_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("compute_streams", [1, 2])
def test_frame_pipeline_matches_direct_calls(compute_streams):
    from transformerupscaler_b200.pipeline import FramePipeline
    M = importlib.import_module("transformerupscaler_b200.models.WindowTransformer.model").TransformerModel().eval()
    M.load_state_dict(synth_state_dict("WindowTransformer", 47), strict=True)
    M = M.to("cuda:0").bfloat16()
    H, W, OH, OW = 72, 104, 108, 156
    frames = [(synth_frames(2, H, W, seed=300 + i) * 255).round().clamp(0, 255).to(torch.uint8) for i in range(7)]
    hin = [f.pin_memory() for f in frames]
    hout = [torch.zeros((2, 3, OH, OW), dtype=torch.uint8).pin_memory() for _ in frames]
    pipe = FramePipeline(M, depth=3, device=torch.device("cuda:0"), compute_streams=compute_streams, res_out=(OH, OW))
    for a, b in zip(hin, hout):
        pipe.submit(a, b)
    pipe.drain()
    with torch.no_grad():
        for a, b in zip(frames, hout):
            ref = M(a.to("cuda:0"), res_out=(OH, OW))
            assert ref.dtype == torch.uint8 and tuple(ref.shape) == (2, 3, OH, OW)
            assert torch.equal(b, ref.cpu())
