"""ORACLE (test infrastructure, never shipped): fp32 torch restatement of AlphaZeroNet.forward.

Functional form of /root/reference/alpha_zero/core/network.py:85-173 driven directly by a state_dict
(so it needs neither the reference package nor the shipped module): conv3x3(pad 3 for Gomoku, :101)
-> BN -> ReLU, residual blocks (:42-82), policy head (:127-139), value head (:141-156).
This is the "plain torch fp32 reference" the CUDA network kernels are compared with.
"""
import torch
import torch.nn.functional as F


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + '.running_mean'], sd[prefix + '.running_var'], sd[prefix + '.weight'], sd[prefix + '.bias'], False, 0.0, 1e-5)


def num_res_blocks(sd):
    return 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('res_blocks.'))


@torch.no_grad()
def forward(sd, x, gomoku):
    """x: float32 [B,17,N,N] -> (pi_logits [B,A], value [B,1])."""
    sd = {k: (torch.as_tensor(v)) for k, v in sd.items()}
    h = F.relu(_bn(F.conv2d(x, sd['conv_block.0.weight'], padding=3 if gomoku else 1), sd, 'conv_block.1'))
    for i in range(num_res_blocks(sd)):
        p = f'res_blocks.{i}.'
        t = F.relu(_bn(F.conv2d(h, sd[p + 'conv_block1.0.weight'], padding=1), sd, p + 'conv_block1.1'))
        t = _bn(F.conv2d(t, sd[p + 'conv_block2.0.weight'], padding=1), sd, p + 'conv_block2.1')
        h = F.relu(t + h)
    pol = F.relu(_bn(F.conv2d(h, sd['policy_head.0.weight']), sd, 'policy_head.1')).flatten(1)
    logits = F.linear(pol, sd['policy_head.4.weight'], sd['policy_head.4.bias'])
    val = F.relu(_bn(F.conv2d(h, sd['value_head.0.weight']), sd, 'value_head.1')).flatten(1)
    val = F.relu(F.linear(val, sd['value_head.4.weight'], sd['value_head.4.bias']))
    val = torch.tanh(F.linear(val, sd['value_head.6.weight'], sd['value_head.6.bias']))
    return logits, val


def make_eval_func(sd, gomoku):
    """eval_position (pipeline.py:91-123): int8 obs -> (softmax priors as float32 arrays, Python-float values)."""

    def eval_func(obs, batched=False):
        import numpy as np

        st = obs if batched else obs[None, ...]
        logits, v = forward(sd, torch.from_numpy(np.ascontiguousarray(st)).to(torch.float32), gomoku)
        pi = torch.softmax(logits, dim=-1).numpy()
        v = v.numpy().squeeze(1).tolist()
        pis = [pi[i] for i in range(pi.shape[0])]
        return (pis, v) if batched else (pis[0], v[0])

    return eval_func


@torch.no_grad()
def forward_bf16_emulated(sd, x, gomoku):
    """What the tensor-core tower computes, restated in torch: BN folded in fp32, conv weights and every
    stored activation rounded to bfloat16, accumulation / bias / residual / ReLU in fp32, heads in fp32.
    Used to tell kernel bugs from bf16 rounding when checking AZ_NET_BF16 against the fp32 reference."""
    sd = {k: torch.as_tensor(v).float() for k, v in sd.items()}

    def fold(wk, bnk):
        sc = sd[bnk + '.weight'] / torch.sqrt(sd[bnk + '.running_var'] + 1e-5)
        return sd[wk] * sc.view(-1, 1, 1, 1), sd[bnk + '.bias'] - sd[bnk + '.running_mean'] * sc

    def r(t):
        return t.to(torch.bfloat16).float()

    w, b = fold('conv_block.0.weight', 'conv_block.1')
    h = r(F.relu(F.conv2d(x, r(w), padding=3 if gomoku else 1) + b.view(1, -1, 1, 1)))
    for i in range(num_res_blocks(sd)):
        p = f'res_blocks.{i}.'
        w1, b1 = fold(p + 'conv_block1.0.weight', p + 'conv_block1.1')
        w2, b2 = fold(p + 'conv_block2.0.weight', p + 'conv_block2.1')
        t = r(F.relu(F.conv2d(h, r(w1), padding=1) + b1.view(1, -1, 1, 1)))
        h = r(F.relu(F.conv2d(t, r(w2), padding=1) + b2.view(1, -1, 1, 1) + h))
    wp, bp = fold('policy_head.0.weight', 'policy_head.1')
    pol = F.relu(F.conv2d(h, wp) + bp.view(1, -1, 1, 1)).flatten(1)
    logits = F.linear(pol, sd['policy_head.4.weight'], sd['policy_head.4.bias'])
    wv, bv = fold('value_head.0.weight', 'value_head.1')
    val = F.relu(F.conv2d(h, wv) + bv.view(1, -1, 1, 1)).flatten(1)
    val = F.relu(F.linear(val, sd['value_head.4.weight'], sd['value_head.4.bias']))
    val = torch.tanh(F.linear(val, sd['value_head.6.weight'], sd['value_head.6.bias']))
    return logits, val
