"""VTPModel.enable_cuda_graphs(): captured-graph replays of the inference entry points return exactly what the eager
launches return (same kernels, same order), follow input changes, and are re-captured when the weights change."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=0):
    from oracle.seeded import seeded_state_dict
    from tests.util import load_golden
    from vtp_b200.config import VTPConfig
    from vtp_b200.model import VTPModel

    meta, _ = load_golden("tiny")
    m = VTPModel(VTPConfig(**meta["config"]))
    m.load_state_dict(seeded_state_dict(meta["spec"], seed=seed))
    return m.cuda(), meta


@pytest.mark.parametrize("autocast", [False, True])
def test_graph_replay_equals_eager(autocast):
    from oracle.seeded import seeded_captions, seeded_images, seeded_state_dict

    m, meta = _model()
    x0, x1 = seeded_images(3, 64, 64).cuda(), seeded_images(3, 64, 64, seed=77).cuda()
    ids0, ids1 = seeded_captions(3, 77, 1000).cuda(), seeded_captions(3, 77, 1000, seed=5).cuda()

    def run(x, ids):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            lat = m.get_reconstruction_latents(x)
            return lat, m.get_latents_decoded_images(lat), m.get_clip_image_feature(x), m.get_clip_text_feature(ids)

    eager = [run(x0, ids0), run(x1, ids1)]
    m.enable_cuda_graphs()
    first = run(x0, ids0)                 # builds the four graphs
    assert len(m._graphs) == 4
    again = run(x1, ids1)                 # pure replays on new input values
    back = run(x0, ids0)
    assert len(m._graphs) == 4
    for got, want in ((first, eager[0]), (again, eager[1]), (back, eager[0])):
        for a, b in zip(got, want):
            assert a.dtype == b.dtype and torch.equal(a, b)
    assert first[0].data_ptr() != back[0].data_ptr()          # outputs are fresh tensors, not the static buffer
    # another batch size -> its own graphs; weight change -> re-capture with the new weights
    small = m.get_reconstruction_latents(x0[:1])
    m.enable_cuda_graphs(False)
    assert torch.equal(small, m.get_reconstruction_latents(x0[:1]))
    m.enable_cuda_graphs()
    m.get_reconstruction_latents(x0)
    m.load_state_dict(seeded_state_dict(meta["spec"], seed=1))
    new_graph = m.get_reconstruction_latents(x0)
    m.enable_cuda_graphs(False)
    assert torch.equal(new_graph, m.get_reconstruction_latents(x0))
    assert not torch.equal(new_graph, eager[0][0].to(new_graph.dtype))
