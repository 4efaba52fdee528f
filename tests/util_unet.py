"""Test helper: an all-zeros reference-layout state_dict derived from the oracle's plan
(key grammar of SURVEY App. A.3; key ORDER follows nn.Module registration order of the
reference so the layout hash can be compared)."""
import torch

from oracle import unet_oracle as UO


def empty_state_dict(img_size, base_channels, attention_resolutions="32,16,8", in_channels=1, channel_mults="",
                     num_res_blocks=2):
    down, middle, up = UO.plan(img_size, base_channels, channel_mults, num_res_blocks, attention_resolutions)
    E = 4 * base_channels
    sd = {}

    def z(*s):
        return torch.zeros(*s)

    sd["time_embedding.1.weight"], sd["time_embedding.1.bias"] = z(E, base_channels), z(E)
    sd["time_embedding.3.weight"], sd["time_embedding.3.bias"] = z(E, E), z(E)

    def block(prefix, blk):
        for j, e in enumerate(blk):
            p = f"{prefix}.{j}"
            if e == "stem":
                sd[p + ".weight"], sd[p + ".bias"] = z(base_channels, in_channels, 3, 3), z(base_channels)
            elif e[0] == "res":
                _, cin, cout, _ = e
                sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"] = z(cin), z(cin)
                sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"] = z(cout, cin, 3, 3), z(cout)
                sd[p + ".embed_layers.1.weight"], sd[p + ".embed_layers.1.bias"] = z(cout, E), z(cout)
                sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"] = z(cout), z(cout)
                sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"] = z(cout, cout, 3, 3), z(cout)
                if cin != cout:
                    sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"] = z(cout, cin, 1, 1), z(cout)
            else:
                c = e[1]
                sd[p + ".norm.weight"], sd[p + ".norm.bias"] = z(c), z(c)
                sd[p + ".to_qkv.weight"], sd[p + ".to_qkv.bias"] = z(3 * c, c, 1), z(3 * c)
                sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"] = z(c, c, 1), z(c)

    for i, blk in enumerate(down):
        block(f"down.{i}", blk)
    block("middle", middle)
    for i, blk in enumerate(up):
        block(f"up.{i}", blk)
    c0 = up[-1][0][2]
    sd["out.0.weight"], sd["out.0.bias"] = z(c0), z(c0)
    sd["out.2.weight"], sd["out.2.bias"] = z(in_channels, c0, 3, 3), z(in_channels)
    return sd
