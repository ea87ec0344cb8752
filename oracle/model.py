"""SSD300 / SSD512 / SSD7 forward graphs (oracle, torch-CPU float32).  Not checkable against Keras itself (not installable
offline); PINNED against the outputs of the reference's real builders executed over eager stand-in Keras layers
(tests/golden/make_tf_golden.py, tests/test_oracle_tf_shim_golden.py).

TensorFlow 1.x / Keras 2.x cannot be installed offline, so these restate the graphs from
  * models/keras_ssd300.py:263-419
  * models/keras_ssd512.py:263-439 (differs from 300 at :43,47,176,312-321,336,345,368-370)
  * models/keras_ssd7.py:266-393
  * keras_layers/keras_layer_L2Normalization.py:61-63 (+ tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12)))
with torch.nn.functional on CPU.  Conventions restated: Keras 'same' conv = symmetric pad for odd
kernels; TF 'same' max-pool on an odd extent pads at the END (== ceil_mode=True for 2x2/s2);
pool5 is 3x3/s1 pad 1; fc6 dilation 6 pad 6; ZeroPadding2D(1)+'valid' == padding=1;
Keras BatchNormalization eps=1e-3; ELU alpha=1; kernels are HWIO; prediction row layout
[softmax(C) | 4 offsets | 4 anchor (cx,cy,w,h) | 4 variances]; prior index ((y*W+x)*n_boxes+b),
head channel b*C+c.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

from .anchors import all_anchors, boxes_per_cell, resolve_scales

VGG_CFG = [  # name, cout, k, dilation
    ('conv1_1', 64), ('conv1_2', 64), 'pool',
    ('conv2_1', 128), ('conv2_2', 128), 'pool',
    ('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256), 'pool',
    ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512), 'pool',
    ('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512),
]

SSD300_AR = [[1.0, 2.0, 0.5], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0],
             [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5], [1.0, 2.0, 0.5]]
SSD512_AR = [[1.0, 2.0, 0.5], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0],
             [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5, 3.0, 1.0 / 3.0], [1.0, 2.0, 0.5], [1.0, 2.0, 0.5]]


def _conv(x, w_hwio, b, stride=1, padding=0, dilation=1):
    w = torch.as_tensor(np.ascontiguousarray(np.transpose(w_hwio, (3, 2, 0, 1))))
    return Fn.conv2d(x, w, torch.as_tensor(b), stride=stride, padding=padding, dilation=dilation)


def _preprocess(x_nhwc, subtract_mean, divide_by_stddev, swap_channels):
    """models/keras_ssd300.py:247-272: mean -> stddev -> channel swap, in that order."""
    x = torch.as_tensor(np.asarray(x_nhwc, dtype=np.float32))
    if subtract_mean is not None:
        x = x - torch.tensor(subtract_mean, dtype=torch.float32)
    if divide_by_stddev is not None:
        x = x / torch.tensor(divide_by_stddev, dtype=torch.float32)
    if swap_channels:
        x = x[..., list(swap_channels)]
    return x.permute(0, 3, 1, 2).contiguous()


def l2_normalize(x_nchw, gamma):
    """keras_layer_L2Normalization.py:61-63."""
    ss = torch.sum(x_nchw * x_nchw, dim=1, keepdim=True)
    return x_nchw * torch.rsqrt(torch.clamp(ss, min=1e-12)) * torch.as_tensor(gamma).view(1, -1, 1, 1)


def _assemble(sources, weights, head_names, n_classes_total, anchors, variances, logits_out=None):
    """Reshape/Concat/softmax/Concat, models/keras_ssd300.py:363-419."""
    confs, locs = [], []
    for src, (cname, lname) in zip(sources, head_names):
        c = _conv(src, weights[cname + '/kernel'], weights[cname + '/bias'], padding=1)
        l = _conv(src, weights[lname + '/kernel'], weights[lname + '/bias'], padding=1)
        B = c.shape[0]
        confs.append(c.permute(0, 2, 3, 1).reshape(B, -1, n_classes_total))
        locs.append(l.permute(0, 2, 3, 1).reshape(B, -1, 4))
    logits = torch.cat(confs, dim=1)
    if logits_out is not None:
        logits_out['logits'] = logits.numpy()
    conf = torch.softmax(logits, dim=-1)
    loc = torch.cat(locs, dim=1)
    B, P = conf.shape[0], conf.shape[1]
    anc = torch.as_tensor(anchors.astype(np.float32)).unsqueeze(0).expand(B, P, 4)
    var = torch.as_tensor(np.asarray(variances, dtype=np.float32)).view(1, 1, 4).expand(B, P, 4)
    return torch.cat([conf, loc, anc, var], dim=-1).numpy()


def ssd_vgg_forward(x_nhwc, weights, variant=300, n_classes=20, scales=None, min_scale=None, max_scale=None,
                    aspect_ratios_per_layer=None, two_boxes_for_ar1=True, steps=None, offsets=None,
                    clip_boxes=False, variances=(0.1, 0.1, 0.2, 0.2), coords='centroids', normalize_coords=True,
                    subtract_mean=(123, 117, 104), divide_by_stddev=None, swap_channels=(2, 1, 0),
                    return_features=False):
    """SSD300 (variant=300) or SSD512 (variant=512), mode='training' output (B,P,C+12)."""
    C = n_classes + 1
    if aspect_ratios_per_layer is None:
        aspect_ratios_per_layer = SSD300_AR if variant == 300 else SSD512_AR
    if steps is None:
        steps = [8, 16, 32, 64, 100, 300] if variant == 300 else [8, 16, 32, 64, 128, 256, 512]
    img_h, img_w = x_nhwc.shape[1], x_nhwc.shape[2]
    x = _preprocess(x_nhwc, subtract_mean, divide_by_stddev, swap_channels)
    feats = {}
    n_pool = 0
    for item in VGG_CFG:
        if item == 'pool':
            n_pool += 1
            x = Fn.max_pool2d(x, 2, 2, ceil_mode=True)
        else:
            name = item[0]
            x = torch.relu(_conv(x, weights[name + '/kernel'], weights[name + '/bias'], padding=1))
            feats[name] = x
    x = Fn.max_pool2d(x, 3, 1, padding=1)                                        # pool5
    x = torch.relu(_conv(x, weights['fc6/kernel'], weights['fc6/bias'], padding=6, dilation=6)); feats['fc6'] = x
    x = torch.relu(_conv(x, weights['fc7/kernel'], weights['fc7/bias'])); feats['fc7'] = x

    def extra(x, n1, n2, stride, pad, k=3):
        x = torch.relu(_conv(x, weights[n1 + '/kernel'], weights[n1 + '/bias']))
        x = torch.relu(_conv(x, weights[n2 + '/kernel'], weights[n2 + '/bias'], stride=stride, padding=pad))
        feats[n2] = x
        return x
    x = extra(x, 'conv6_1', 'conv6_2', 2, 1)
    x = extra(x, 'conv7_1', 'conv7_2', 2, 1)
    if variant == 300:
        x = extra(x, 'conv8_1', 'conv8_2', 1, 0)
        x = extra(x, 'conv9_1', 'conv9_2', 1, 0)
        src_names = ['conv4_3_norm', 'fc7', 'conv6_2', 'conv7_2', 'conv8_2', 'conv9_2']
    else:
        x = extra(x, 'conv8_1', 'conv8_2', 2, 1)
        x = extra(x, 'conv9_1', 'conv9_2', 2, 1)
        x = extra(x, 'conv10_1', 'conv10_2', 1, 1, k=4)                          # 4x4 valid after pad 1
        src_names = ['conv4_3_norm', 'fc7', 'conv6_2', 'conv7_2', 'conv8_2', 'conv9_2', 'conv10_2']
    feats['conv4_3_norm'] = l2_normalize(feats['conv4_3'], weights['conv4_3_norm/gamma'])
    sources = [feats[n] for n in src_names]
    sizes = [(s.shape[2], s.shape[3]) for s in sources]
    sc = resolve_scales(len(sizes), min_scale, max_scale, scales)
    anchors = all_anchors(img_h, img_w, sizes, sc, aspect_ratios_per_layer, two_boxes_for_ar1, steps, offsets,
                          clip_boxes, coords, normalize_coords)
    heads = [(n + '_mbox_conf', n + '_mbox_loc') for n in src_names]
    extra_out = {}
    y = _assemble(sources, weights, heads, C, anchors, variances, extra_out)
    if return_features:
        f = {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in feats.items()}
        f.update(extra_out)
        return y, f
    return y


def ssd7_forward(x_nhwc, weights, n_classes=5, min_scale=0.1, max_scale=0.9, scales=None,
                 aspect_ratios_global=(0.5, 1.0, 2.0), aspect_ratios_per_layer=None, two_boxes_for_ar1=True,
                 steps=None, offsets=None, clip_boxes=False, variances=(1.0, 1.0, 1.0, 1.0), coords='centroids',
                 normalize_coords=False, subtract_mean=None, divide_by_stddev=None, swap_channels=False,
                 return_features=False):
    """build_model, models/keras_ssd7.py:266-393, mode='training' output."""
    C = n_classes + 1
    img_h, img_w = x_nhwc.shape[1], x_nhwc.shape[2]
    x = _preprocess(x_nhwc, subtract_mean, divide_by_stddev, swap_channels)
    feats = {}
    for i in range(1, 8):
        k = weights['conv%d/kernel' % i].shape[0]
        x = _conv(x, weights['conv%d/kernel' % i], weights['conv%d/bias' % i], padding=k // 2)
        g, b_, m, v = (weights['bn%d/%s' % (i, s)] for s in ('gamma', 'beta', 'moving_mean', 'moving_variance'))
        x = Fn.batch_norm(x, torch.as_tensor(m), torch.as_tensor(v), torch.as_tensor(g), torch.as_tensor(b_),
                          training=False, eps=1e-3)
        x = Fn.elu(x)
        feats['conv%d' % i] = x
        if i < 7:
            x = Fn.max_pool2d(x, 2, 2)                                           # Keras default: valid
    sources = [feats['conv4'], feats['conv5'], feats['conv6'], feats['conv7']]
    sizes = [(s.shape[2], s.shape[3]) for s in sources]
    sc = resolve_scales(4, min_scale, max_scale, scales)
    ar = aspect_ratios_per_layer if aspect_ratios_per_layer is not None else [list(aspect_ratios_global)] * 4
    anchors = all_anchors(img_h, img_w, sizes, sc, ar, two_boxes_for_ar1, steps, offsets, clip_boxes, coords,
                          normalize_coords)
    heads = [('classes%d' % i, 'boxes%d' % i) for i in (4, 5, 6, 7)]
    extra_out = {}
    y = _assemble(sources, weights, heads, C, anchors, variances, extra_out)
    if return_features:
        f = {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in feats.items()}
        f.update(extra_out)
        return y, f
    return y


# ---------------------------------------------------------------------------------------------
# weight shapes (HWIO) so tests/bench can synthesise he_normal weights for both implementations
# ---------------------------------------------------------------------------------------------

def vgg_weight_shapes(variant=300, n_classes=20, aspect_ratios_per_layer=None, two_boxes_for_ar1=True):
    C = n_classes + 1
    ar = aspect_ratios_per_layer or (SSD300_AR if variant == 300 else SSD512_AR)
    nb = [boxes_per_cell(a, two_boxes_for_ar1) for a in ar]
    shapes = {}
    cin = 3
    for item in VGG_CFG:
        if item == 'pool':
            continue
        shapes[item[0]] = (3, 3, cin, item[1]); cin = item[1]
    shapes['fc6'] = (3, 3, 512, 1024)
    shapes['fc7'] = (1, 1, 1024, 1024)
    shapes['conv6_1'] = (1, 1, 1024, 256); shapes['conv6_2'] = (3, 3, 256, 512)
    shapes['conv7_1'] = (1, 1, 512, 128); shapes['conv7_2'] = (3, 3, 128, 256)
    shapes['conv8_1'] = (1, 1, 256, 128); shapes['conv8_2'] = (3, 3, 128, 256)
    shapes['conv9_1'] = (1, 1, 256, 128); shapes['conv9_2'] = (3, 3, 128, 256)
    srcs = [('conv4_3_norm', 512), ('fc7', 1024), ('conv6_2', 512), ('conv7_2', 256), ('conv8_2', 256), ('conv9_2', 256)]
    if variant == 512:
        shapes['conv10_1'] = (1, 1, 256, 128); shapes['conv10_2'] = (4, 4, 128, 256)
        srcs.append(('conv10_2', 256))
    for (n, c), b in zip(srcs, nb):
        shapes[n + '_mbox_conf'] = (3, 3, c, b * C)
        shapes[n + '_mbox_loc'] = (3, 3, c, b * 4)
    return shapes


def ssd7_weight_shapes(n_classes=5, n_boxes=(4, 4, 4, 4)):
    C = n_classes + 1
    chans = [3, 32, 48, 64, 64, 48, 48, 32]
    shapes = {'conv1': (5, 5, 3, 32)}
    for i in range(2, 8):
        shapes['conv%d' % i] = (3, 3, chans[i - 1], chans[i])
    for j, i in enumerate((4, 5, 6, 7)):
        shapes['classes%d' % i] = (3, 3, chans[i], n_boxes[j] * C)
        shapes['boxes%d' % i] = (3, 3, chans[i], n_boxes[j] * 4)
    return shapes
