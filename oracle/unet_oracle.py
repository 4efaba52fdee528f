"""Plain-PyTorch fp32 functional restatement of the reference UNet forward
(TEST INFRASTRUCTURE; see oracle/__init__.py).  The UNet is floating-point work, so this
keeps a torch fp32 reference: every op is a stock torch.nn.functional call on CPU (or on
the GPU when used as the "library comparator").

Works directly on a reference-layout ``state_dict`` (keys of SURVEY App. A.3); pinned
against the unmodified reference UNetModel by tests/golden/unet_*.npz.

Reference lines followed (all /root/reference/UNet.py):
  plan()              <- UNetModel.__init__        :222-388 (which block sits where)
  timestep_embedding  <- PositionalEmbedding       :38-57, time MLP :271-276
  resblock            <- ResBlock.forward          :202-217
  attention           <- AttentionBlock/QKVAttention :119-153
  forward             <- UNetModel.forward         :390-406
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def default_mults(img_size):
    table = {512: (0.5, 1, 1, 2, 2, 4, 4), 256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4),
             32: (1, 2, 3, 4)}
    if img_size not in table:
        raise ValueError(f"unsupported image size: {img_size}")
    return table[img_size]


def plan(img_size, base_channels, channel_mults="", num_res_blocks=2, attention_resolutions="32,16,8"):
    """Returns (down, middle, up): lists of entries; each entry is a list of
    ("res", cin, cout, mode) / ("attn", c) tuples, mode in {"same","down","up"}."""
    if channel_mults == "":
        channel_mults = default_mults(img_size)
    attn_ds = [img_size // int(r) for r in attention_resolutions.split(",")]
    ch = int(channel_mults[0] * base_channels)
    down, stack, ds = [["stem"]], [ch], 1
    for lvl, m in enumerate(channel_mults):
        for _ in range(num_res_blocks):
            blk = [("res", ch, base_channels * m, "same")]
            ch = base_channels * m
            if ds in attn_ds:
                blk.append(("attn", ch))
            down.append(blk)
            stack.append(ch)
        if lvl != len(channel_mults) - 1:
            down.append([("res", ch, ch, "down")])
            ds *= 2
            stack.append(ch)
    middle = [("res", ch, ch, "same"), ("attn", ch), ("res", ch, ch, "same")]
    up = []
    for lvl, m in reversed(list(enumerate(channel_mults))):
        for j in range(num_res_blocks + 1):
            blk = [("res", ch + stack.pop(), base_channels * m, "same")]
            ch = base_channels * m
            if ds in attn_ds:
                blk.append(("attn", ch))
            if lvl and j == num_res_blocks:
                blk.append(("res", ch, ch, "up"))
                ds //= 2
            up.append(blk)
    return down, middle, up


def timestep_embedding(sd, t, base_channels):
    half = base_channels // 2
    w = torch.exp(torch.arange(half, device=t.device) * -(np.log(10000) / half))
    e = torch.outer(t * 1, w)
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = F.linear(e, sd["time_embedding.1.weight"], sd["time_embedding.1.bias"])
    return F.linear(F.silu(e), sd["time_embedding.3.weight"], sd["time_embedding.3.bias"])


def _gn(sd, key, x):
    return F.group_norm(x.float(), 32, sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def resblock(sd, p, x, emb, mode):
    h = F.silu(_gn(sd, p + ".in_layers.0", x))
    if mode == "down":
        h, x = F.avg_pool2d(h, 2, 2), F.avg_pool2d(x, 2, 2)
    elif mode == "up":
        h, x = F.interpolate(h, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".embed_layers.1.weight"], sd[p + ".embed_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.silu(_gn(sd, p + ".out_layers.0", h))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        w = sd[p + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def attention(sd, p, x, n_heads, n_head_channels):
    b, c, hh, ww = x.shape
    heads = n_heads if n_head_channels == -1 else c // n_head_channels
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + ".norm", xf), sd[p + ".to_qkv.weight"], sd[p + ".to_qkv.bias"])
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, 3 * ch, -1).split(ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * s, k * s)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, c, -1)
    a = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + a).reshape(b, c, hh, ww)


def forward(sd, x, t, img_size, base_channels, channel_mults="", num_res_blocks=2, attention_resolutions="32,16,8",
            n_heads=1, n_head_channels=-1, taps=None):
    """x (B,C,H,W) fp32, t (B,) int64 -> eps (B,C,H,W).  ``taps`` (dict) collects named intermediates."""
    down, middle, up = plan(img_size, base_channels, channel_mults, num_res_blocks, attention_resolutions)
    emb = timestep_embedding(sd, t, base_channels)

    def run(prefix, blk, h):
        for j, e in enumerate(blk):
            if e == "stem":
                h = F.conv2d(h, sd[f"{prefix}.{j}.weight"], sd[f"{prefix}.{j}.bias"], padding=1)
            elif e[0] == "res":
                h = resblock(sd, f"{prefix}.{j}", h, emb, e[3])
            else:
                h = attention(sd, f"{prefix}.{j}", h, n_heads, n_head_channels)
        return h

    h, skips = x.float(), []
    for i, blk in enumerate(down):
        h = run(f"down.{i}", blk, h)
        skips.append(h)
        if taps is not None:
            taps[f"down.{i}"] = h
    h = run("middle", middle, h)
    if taps is not None:
        taps["middle"] = h
    for i, blk in enumerate(up):
        h = run(f"up.{i}", blk, torch.cat([h, skips.pop()], dim=1))
        if taps is not None:
            taps[f"up.{i}"] = h
    h = F.silu(_gn(sd, "out.0", h))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def fill_state_dict(sd, seed=0):
    """Deterministic, key-addressed synthetic weights: the filler lives in anoddpm_b200/synth.py (it is input generation,
    not oracle arithmetic, and bench.py's native arm must not import oracle/); kept here under its old name for the tests."""
    from anoddpm_b200.synth import fill_state_dict as _fill
    return _fill(sd, seed)
