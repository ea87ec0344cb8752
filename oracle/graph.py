"""Generic torch-CPU executor for a layer-spec graph (oracle; test infrastructure).  Checked against oracle/model.py and
oracle/loss.py (tests/test_oracle_graph_cpu.py), which are pinned to the reference's code (tests/test_oracle_tf_shim_golden.py).

Executes the same ``Spec`` list the product's builders produce (ops INPUT / CONV / MAXPOOL / L2NORM / HEAD) with
torch.nn.functional in float32 or float64, keeping the autograd graph so that gradients of the SSD loss with respect to
every kernel / bias / gamma can be compared with the CUDA backward pass.  Conventions as in oracle/model.py.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

OP_INPUT, OP_CONV, OP_MAXPOOL, OP_L2NORM, OP_HEAD = range(5)
ACT_NONE, ACT_RELU, ACT_ELU = range(3)


def make_params(specs, weights, dtype=torch.float32, requires_grad=True):
    """numpy weight dict (Keras names) -> dict of torch leaf tensors."""
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad) for k, v in weights.items()}


def forward(specs, params, x_nhwc, n_classes_total, anchors, variances, dtype=torch.float32, bn_training=False):
    """-> y_pred (B, P, C+12) torch tensor (differentiable w.r.t. ``params``).  ``bn_training``: BatchNormalization in Keras'
    training phase (what ``fit_generator`` runs, models/keras_ssd7.py:277-309): batch mean / biased variance over (B,H,W), epsilon
    1e-3; the batch statistics are returned in ``outs['<bn>/batch_mean' | '/batch_var']`` (biased variance)."""
    outs = {}
    confs, locs = [], []
    for s in specs:
        if s.op == OP_INPUT:
            x = torch.as_tensor(np.asarray(x_nhwc), dtype=dtype)
            p = s.params
            if p.get('mean') is not None:
                x = x - torch.tensor(np.asarray(p['mean'], dtype=np.float64), dtype=dtype)
            if p.get('stddev') is not None:
                x = x / torch.tensor(np.asarray(p['stddev'], dtype=np.float64), dtype=dtype)
            if p.get('swap'):
                x = x[..., list(p['swap'])]
            outs[s.name] = x.permute(0, 3, 1, 2).contiguous()
            continue
        xin = outs[s.inp]
        pt, pl, pb, pr = s.pad
        if s.op == OP_CONV:
            w = params[s.name + '/kernel'].permute(3, 2, 0, 1)
            y = Fn.conv2d(Fn.pad(xin, (pl, pr, pt, pb)), w, params[s.name + '/bias'], stride=s.stride, dilation=s.dilation)
            if getattr(s, 'bn', None) and bn_training:
                g, b = params[s.bn + '/gamma'], params[s.bn + '/beta']
                mu = y.mean(dim=(0, 2, 3)); var = y.var(dim=(0, 2, 3), unbiased=False)
                outs[s.bn + '/batch_mean'], outs[s.bn + '/batch_var'] = mu.detach(), var.detach()
                y = (y - mu.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-3) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
            elif getattr(s, 'bn', None):        # inference-phase BatchNormalization (Keras epsilon 1e-3) between conv and activation
                g, b = params[s.bn + '/gamma'], params[s.bn + '/beta']
                mu, var = params[s.bn + '/moving_mean'], params[s.bn + '/moving_variance']
                y = (y - mu.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-3) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
            if s.act == ACT_RELU:
                y = torch.relu(y)
            elif s.act == ACT_ELU:
                y = Fn.elu(y)
            outs[s.name] = y
        elif s.op == OP_MAXPOOL:
            xp = Fn.pad(xin, (pl, pr, pt, pb), value=float('-inf'))
            outs[s.name] = Fn.max_pool2d(xp, (s.kh, s.kw), s.stride)
        elif s.op == OP_L2NORM:
            ss = torch.sum(xin * xin, dim=1, keepdim=True)
            outs[s.name] = xin * torch.rsqrt(torch.clamp(ss, min=1e-12)) * params[s.name + '/gamma'].view(1, -1, 1, 1)
        elif s.op == OP_HEAD:
            cn, ln = s.params['conf_name'], s.params['loc_name']
            c = Fn.conv2d(Fn.pad(xin, (pl, pr, pt, pb)), params[cn + '/kernel'].permute(3, 2, 0, 1), params[cn + '/bias'])
            l = Fn.conv2d(Fn.pad(xin, (pl, pr, pt, pb)), params[ln + '/kernel'].permute(3, 2, 0, 1), params[ln + '/bias'])
            B = c.shape[0]
            confs.append(c.permute(0, 2, 3, 1).reshape(B, -1, n_classes_total))
            locs.append(l.permute(0, 2, 3, 1).reshape(B, -1, 4))
    conf = torch.softmax(torch.cat(confs, dim=1), dim=-1)
    loc = torch.cat(locs, dim=1)
    B, P = conf.shape[0], conf.shape[1]
    anc = torch.as_tensor(np.asarray(anchors), dtype=dtype).unsqueeze(0).expand(B, P, 4)
    var = torch.as_tensor(np.asarray(variances), dtype=dtype).view(1, 1, 4).expand(B, P, 4)
    return torch.cat([conf, loc, anc, var], dim=-1), outs


def ssd_loss_torch(y_true, y_pred, neg_pos_ratio=3, n_neg_min=0, alpha=1.0):
    """keras_loss_function/keras_ssd_loss.py:98-211 with torch ops -> (B,) tensor; the hard-negative mask is a constant."""
    yt = torch.as_tensor(np.asarray(y_true), dtype=y_pred.dtype)
    cls = -torch.sum(yt[:, :, :-12] * torch.log(torch.clamp(y_pred[:, :, :-12], min=1e-15)), dim=-1)
    d = yt[:, :, -12:-8] - y_pred[:, :, -12:-8]
    ad = torch.abs(d)
    loc = torch.sum(torch.where(ad < 1.0, 0.5 * d * d, ad - 0.5), dim=-1)
    neg = yt[:, :, 0]
    pos = torch.max(yt[:, :, 1:-12], dim=-1).values
    n_pos = pos.sum()
    B = yt.shape[0]
    neg_all = (cls * neg).detach()
    n_neg_losses = int(torch.count_nonzero(neg_all))
    k = min(max(int(neg_pos_ratio) * int(n_pos.item()), int(n_neg_min)), n_neg_losses)
    mask = torch.zeros_like(neg_all).reshape(-1)
    if k > 0:
        flat = neg_all.reshape(-1).double().numpy()
        order = np.lexsort((np.arange(flat.size), -flat))[:k]
        mask[torch.as_tensor(order)] = 1
    mask = mask.reshape(neg_all.shape)
    total = (torch.sum(cls * pos, -1) + torch.sum(cls * mask, -1) + alpha * torch.sum(loc * pos, -1)) / torch.clamp(n_pos, min=1.0)
    return total * B


def sgd_step(params, grads, velocity, lr, momentum, l2_reg):
    """Keras SGD (v = m*v - lr*g; w += v) with the l2 kernel regulariser's gradient 2*l2*w on kernels.  numpy dicts in/out."""
    new_p, new_v = {}, {}
    for k, w in params.items():
        g = grads[k].astype(np.float64)
        if k.endswith('/kernel'):
            g = g + 2.0 * l2_reg * w.astype(np.float64)
        v = momentum * velocity.get(k, np.zeros_like(w, dtype=np.float64)) - lr * g
        new_v[k] = v
        new_p[k] = (w.astype(np.float64) + v).astype(np.float32)
    return new_p, new_v


def adam_step(params, grads, m, v, t, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-8, l2_reg=0.0):
    """Keras Adam (optimizers.py, decay 0): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    p -= lr_t * m / (sqrt(v) + eps); kernels get the l2 regulariser's gradient 2*l2*w.  ``t`` counts from 1.  numpy dicts."""
    lr_t = lr * np.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)
    new_p, new_m, new_v = {}, {}, {}
    for k, w in params.items():
        if k not in grads:
            new_p[k] = w
            continue
        g = grads[k].astype(np.float64)
        if k.endswith('/kernel'):
            g = g + 2.0 * l2_reg * w.astype(np.float64)
        mk = beta_1 * m.get(k, 0.0) + (1 - beta_1) * g
        vk = beta_2 * v.get(k, 0.0) + (1 - beta_2) * g * g
        new_m[k], new_v[k] = mk, vk
        new_p[k] = (w.astype(np.float64) - lr_t * mk / (np.sqrt(vk) + epsilon)).astype(np.float32)
    return new_p, new_m, new_v
