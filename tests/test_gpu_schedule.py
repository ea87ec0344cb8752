"""GPU tests for the two-stream schedule of inference plans (csrc/model.cu, plan_overlap) and the pipelined host API
(SSDModel.predict_stream / predict_generator / predict).

The schedule only changes WHERE and WHEN the launches run (a second stream for the narrow tail of the trunk and the narrow
predictor heads, a capped persistent grid for the wide heads), never what a work unit computes, so the bar is bit-exact equality
with the single-stream pass of the same plan (the instrumented pass `set_timing(True)` issues everything on one stream with full
grids).  Back-to-back calls with alternating inputs check the cross-call ordering (buffers are reused by the next call while the
side stream may still hold work of the previous one)."""
import numpy as np
import pytest

from oracle import synth
from oracle.model import ssd7_weight_shapes, vgg_weight_shapes

pytestmark = pytest.mark.gpu

SC300 = [0.1, 0.2, 0.37, 0.54, 0.71, 0.88, 1.05]
SC512 = [0.04, 0.1, 0.26, 0.42, 0.58, 0.74, 0.9, 1.06]
PRE = dict(subtract_mean=[123, 117, 104], divide_by_stddev=[64, 64, 64], swap_channels=[2, 1, 0])


@pytest.fixture(scope='module', autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()
    import torch
    assert torch.cuda.is_available()


def _ssd300(mode, B_seed=1, **kw):
    from ssd_keras_b200.models.keras_ssd300 import ssd_300
    m = ssd_300((300, 300, 3), 20, mode=mode, scales=SC300, **PRE, **kw)
    w = synth.synth_weights(B_seed, vgg_weight_shapes(300, 20), bias_scale=0.02)
    w['conv4_3_norm/gamma'] = np.full((512,), 20.0, np.float32)
    m.set_weights(w)
    return m


def _serial_vs_overlap(model, xs, rounds=3):
    """xs: list of device batches of one size.  Overlapped passes back to back (no host sync in between), then the same inputs
    through the single-stream instrumented pass."""
    import torch
    B = xs[0].shape[0]
    got = []
    for r in range(rounds):
        for x in xs:
            got.append(model.forward_device(x).clone())
    torch.cuda.synchronize()
    model.set_timing(B, True)
    ref = [model.forward_device(x).clone() for x in xs]
    torch.cuda.synchronize()
    model.set_timing(B, False)
    for k, g in enumerate(got):
        r = ref[k % len(xs)]
        assert torch.equal(torch.nan_to_num(g, nan=-7.0), torch.nan_to_num(r, nan=-7.0)), 'pass %d differs from the single-stream pass' % k


@pytest.mark.parametrize('B', [32, 4, 1])
def test_ssd300_two_stream_schedule_is_bit_exact(B):
    import torch
    model = _ssd300('training')
    xs = [torch.from_numpy(synth.synth_images(10 + i, B, 300, 300)).cuda() for i in range(2)]
    _serial_vs_overlap(model, xs)


def test_ssd300_two_stream_schedule_forced_reserve_sizes(monkeypatch):
    """Other splits of the SMs between the two streams (SSDK_OVERLAP_R), including one where nearly everything counts as narrow, and
    the schedule switched off."""
    import torch
    xs = [torch.from_numpy(synth.synth_images(20 + i, 8, 300, 300)).cuda() for i in range(2)]
    outs = []
    for env in ({'SSDK_OVERLAP': '0'}, {'SSDK_OVERLAP_R': '8'}, {'SSDK_OVERLAP_R': '64'}, {'SSDK_OVERLAP_R': '120'}):
        for k in ('SSDK_OVERLAP', 'SSDK_OVERLAP_R'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        model = _ssd300('training')                      # the plan reads the knobs when it is created
        _serial_vs_overlap(model, xs, rounds=2)
        outs.append(model.forward_device(xs[0]).clone())
        del model
    for o in outs[1:]:
        assert torch.equal(torch.nan_to_num(o, nan=-7.0), torch.nan_to_num(outs[0], nan=-7.0))


def test_ssd512_and_ssd7_two_stream_schedule():
    import torch
    from ssd_keras_b200.models.keras_ssd512 import ssd_512
    from ssd_keras_b200.models.keras_ssd7 import build_model
    m = ssd_512((512, 512, 3), 80, mode='training', scales=SC512, **PRE)
    w = synth.synth_weights(2, vgg_weight_shapes(512, 80), bias_scale=0.02)
    w['conv4_3_norm/gamma'] = np.full((512,), 20.0, np.float32)
    m.set_weights(w)
    xs = [torch.from_numpy(synth.synth_images(30 + i, 4, 512, 512)).cuda() for i in range(2)]
    _serial_vs_overlap(m, xs, rounds=2)
    del m
    m7 = build_model((300, 300, 3), 5, mode='training', scales=[0.08, 0.16, 0.32, 0.64, 0.96], normalize_coords=True,
                     subtract_mean=[127.5] * 3, divide_by_stddev=[127.5] * 3)
    w = synth.synth_weights(3, ssd7_weight_shapes(5), bias_scale=0.05)
    for i in range(1, 8):
        c = w['conv%d/bias' % i].shape[0]
        w['bn%d/gamma' % i] = np.ones(c, np.float32); w['bn%d/beta' % i] = np.zeros(c, np.float32)
        w['bn%d/moving_mean' % i] = np.zeros(c, np.float32); w['bn%d/moving_variance' % i] = np.ones(c, np.float32)
    m7.set_weights(w)
    xs = [torch.from_numpy(synth.synth_images(40 + i, 16, 300, 300)).cuda() for i in range(2)]
    _serial_vs_overlap(m7, xs, rounds=2)


def test_predict_stream_matches_per_batch_calls():
    """Five different host batches through the pipeline: every result equals the one-batch-at-a-time call, in order; ndarray and
    pinned-tensor inputs; predict_generator with tuples and a step limit; predict with a batch size that does not divide N."""
    import torch
    model = _ssd300('inference')
    B = 4
    host = [synth.synth_images(50 + i, B, 300, 300) for i in range(5)]
    ref = []
    for h in host:
        ref.append(model.predict_device(torch.from_numpy(h).cuda()).cpu().numpy())
    torch.cuda.synchronize()
    got = [r.numpy().copy() for r in model.predict_stream(iter(host))]
    assert len(got) == 5
    for g, r in zip(got, ref):
        assert np.array_equal(np.nan_to_num(g, nan=-7.0), np.nan_to_num(r, nan=-7.0))
    pinned = [torch.from_numpy(h).pin_memory() for h in host]
    got2 = [r.numpy().copy() for r in model.predict_stream(pinned)]
    for g, r in zip(got2, ref):
        assert np.array_equal(np.nan_to_num(g, nan=-7.0), np.nan_to_num(r, nan=-7.0))
    # Keras-style generator of (X, y) tuples, limited to 3 steps
    out = model.predict_generator(((h, None) for h in host), steps=3)
    assert out.shape == (3 * B, 200, 6)
    assert np.array_equal(np.nan_to_num(out, nan=-7.0), np.nan_to_num(np.concatenate(ref[:3]), nan=-7.0))
    # predict(): 10 images in batches of 4 (4 + 4 + 2): per-image results do not depend on the batch they travel in
    x = np.concatenate(host)[:10]
    y = model.predict(x, batch_size=4)
    assert y.shape == (10, 200, 6)
    y1 = model.predict(x[8:10])                        # (a plan for batch 2 may tile differently: class ids exact, values close)
    assert np.array_equal(y[8:10, :, 0], y1[:, :, 0])
    assert np.allclose(np.nan_to_num(y[8:10], nan=-7.0), np.nan_to_num(y1, nan=-7.0), rtol=1e-4, atol=1e-4)
    assert np.array_equal(np.nan_to_num(y[:8], nan=-7.0), np.nan_to_num(np.concatenate(ref[:2]), nan=-7.0))
    # an empty generator
    assert list(model.predict_stream(iter([]))) == []
    with pytest.raises(ValueError):
        model.predict_generator(iter([]))
