"""GPU: the host-buffer frame pipeline (transformerupscaler_b200.pipeline.FramePipeline: copy-in / copy-out streams and alternating
compute streams, the path bench.py's `e2e` number goes through) returns, batch for batch, exactly the bytes of a plain synchronous
call of the same model -- also when the forwards of consecutive batches overlap on two streams.  (Replaces the synchronous
.to(device) -> model -> .cpu() loops of speed_test.py:60-67 / app_overlay.py:365-391.)"""
import importlib

import pytest
import torch

from oracle.weights import synth_state_dict, synth_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("compute_streams,depth", [(1, 3), (2, 3), (3, 4)])      # (3, 4): what bench.py's e2e region runs
def test_frame_pipeline_matches_direct_calls(compute_streams, depth):
    from transformerupscaler_b200.pipeline import FramePipeline
    M = importlib.import_module("transformerupscaler_b200.models.WindowTransformer.model").TransformerModel().eval()
    M.load_state_dict(synth_state_dict("WindowTransformer", 47), strict=True)
    M = M.to("cuda:0").bfloat16()
    H, W, OH, OW = 72, 104, 108, 156
    frames = [(synth_frames(2, H, W, seed=300 + i) * 255).round().clamp(0, 255).to(torch.uint8) for i in range(9)]
    hin = [f.pin_memory() for f in frames]
    hout = [torch.zeros((2, 3, OH, OW), dtype=torch.uint8).pin_memory() for _ in frames]
    pipe = FramePipeline(M, depth=depth, device=torch.device("cuda:0"), compute_streams=compute_streams, res_out=(OH, OW))
    for a, b in zip(hin, hout):
        pipe.submit(a, b)
    pipe.drain()
    with torch.no_grad():
        for a, b in zip(frames, hout):
            ref = M(a.to("cuda:0"), res_out=(OH, OW))
            assert ref.dtype == torch.uint8 and tuple(ref.shape) == (2, 3, OH, OW)
            assert torch.equal(b, ref.cpu())


def test_pipeline_tickets_allow_a_ring_of_pinned_buffers():
    """A video caller recycles a small ring of pinned buffers: the Ticket says when an input buffer has been consumed and when an
    output buffer is complete, so overwriting the ring does not corrupt frames in flight."""
    from transformerupscaler_b200.pipeline import FramePipeline
    M = importlib.import_module("transformerupscaler_b200.models.WindowTransformer.model").TransformerModel().eval()
    M.load_state_dict(synth_state_dict("WindowTransformer", 48), strict=True)
    M = M.to("cuda:0").bfloat16()
    H, W, OH, OW = 72, 104, 108, 156
    frames = [(synth_frames(2, H, W, seed=400 + i) * 255).round().clamp(0, 255).to(torch.uint8) for i in range(9)]
    ring_in = [torch.empty((2, 3, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    ring_out = [torch.empty((2, 3, OH, OW), dtype=torch.uint8).pin_memory() for _ in range(2)]
    pipe = FramePipeline(M, depth=2, device=torch.device("cuda:0"), compute_streams=2, res_out=(OH, OW))
    tickets, got = [None, None], []
    for i, f in enumerate(frames):
        k = i % 2
        if tickets[k] is not None:
            tickets[k].wait()                      # the ring's output buffer is complete: consume it before it is rewritten
            got.append(ring_out[k].clone())
            assert tickets[k].input_consumed() and tickets[k].done()
        ring_in[k].copy_(f)                        # safe: the previous user of this buffer has been waited for
        tickets[k] = pipe.submit(ring_in[k], ring_out[k])
    for j in range(len(frames) - 2, len(frames)):
        tickets[j % 2].wait()
        got.append(ring_out[j % 2].clone())
    with torch.no_grad():
        for f, o in zip(frames, got):
            assert torch.equal(o, M(f.to("cuda:0"), res_out=(OH, OW)).cpu())


@pytest.mark.parametrize("model,shape,kw", [("WindowTransformer", (1, 3, 720, 1280), dict(res_out=(1080, 1920))),
                                            ("FastTransformer", (1, 3, 96, 128), dict(upscale_factor=2)),
                                            ("ResidualTransformer", (1, 3, 720, 1280), dict(res_out=(1080, 1920)))])
def test_cuda_graph_replay_is_bitwise_the_eager_forward(model, shape, kw):
    """transformerupscaler_b200.graph.GraphedModel: the whole forward captured into a CUDA graph (programmatic-launch edges, tile
    flags and workspace included) and replayed on new frames gives exactly the eager result."""
    from transformerupscaler_b200.graph import GraphedModel
    M = importlib.import_module(f"transformerupscaler_b200.models.{model}.model").TransformerModel().eval()
    M.load_state_dict(synth_state_dict(model, 49), strict=True)
    M = M.to("cuda:0").bfloat16()
    G = GraphedModel(M)
    for i in range(3):
        x = synth_frames(shape[0], shape[2], shape[3], seed=500 + i).cuda().bfloat16()
        with torch.no_grad():
            want = M(x, **kw)
        got = G(x, **kw)
        assert torch.equal(got, want), f"replay {i} differs"
    assert len(G._graphs) == 1
