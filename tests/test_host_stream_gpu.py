"""GPU: pipeline.HostStream (pipelined pinned-host -> pinned-host evaluation) returns, batch by batch, exactly what the
synchronous predict_device path returns, including a ragged last batch and slot reuse."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_host_stream_matches_synchronous_path():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import pipeline, weights
    S, B, seed = 224, 3, 2
    sd = weights.make_visual_state_dict(seed=seed, img_size=S)
    pred = pipeline.AnomalyPredictor(sd, weights.make_text_anchor(seed), img_size=S, device="cuda:0", max_batch=B)
    batches = [weights.make_images(seed, b, S, first_index=10 * i).pin_memory() for i, b in enumerate([3, 3, 2, 3, 1])]
    outs = [(torch.empty((x.shape[0], S, S)).pin_memory(), torch.empty((x.shape[0],)).pin_memory()) for x in batches]
    hs = pipeline.HostStream(pred, slots=2)
    events = [hs.submit(x, m, s) for x, (m, s) in zip(batches, outs)]
    for ev in events:
        ev.synchronize()
    hs.drain()
    for x, (m, s) in zip(batches, outs):
        maps, scores = pred.predict_device(x.cuda())
        torch.cuda.synchronize()
        assert torch.equal(maps.cpu(), m) and torch.equal(scores.cpu(), s)
