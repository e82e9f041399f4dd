"""Static description of the NCSN++ vector-field backbone (default FlowSE config).

The reference builds a flat ``all_modules`` list in its constructor
(/root/reference/flowmse/backbones/ncsnpp.py:99-245) and walks it by index in
``forward`` (:247-404).  This module re-derives that list as plain data so that
the checkpoint reader, the parameter-holding ``NCSNpp`` shim and the C-ABI weight
packer all agree on names, shapes and order (647 tensors, state-dict order).
"""
from __future__ import annotations

from typing import List, Tuple

NF = 128
CH_MULT = (1, 1, 2, 2, 2, 2, 2)
NUM_RES_BLOCKS = 2
ATTN_RESOLUTIONS = (16,)
IMAGE_SIZE = 256           # F must be 256 so that attention lands on H == 16 (ncsnpp.py:296,342)
NUM_CHANNELS = 4           # x.re, x.im, y.re, y.im (ncsnpp.py:95)
TEMB_DIM = 4 * NF
FOURIER_SCALE = 16.0


def module_list() -> List[dict]:
    """The 77 entries of ``all_modules``."""
    mods: List[dict] = []
    num_res = len(CH_MULT)
    all_res = [IMAGE_SIZE // (2 ** i) for i in range(num_res)]
    mods.append(dict(kind="fourier", size=NF))
    mods.append(dict(kind="linear", cin=2 * NF, cout=TEMB_DIM))
    mods.append(dict(kind="linear", cin=TEMB_DIM, cout=TEMB_DIM))
    mods.append(dict(kind="conv3x3", cin=NUM_CHANNELS, cout=NF))
    hs_c = [NF]
    in_ch = NF
    for lvl in range(num_res):
        for _ in range(NUM_RES_BLOCKS):
            out_ch = NF * CH_MULT[lvl]
            mods.append(dict(kind="rb", cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            if all_res[lvl] in ATTN_RESOLUTIONS:
                mods.append(dict(kind="attn", c=in_ch))
            hs_c.append(in_ch)
        if lvl != num_res - 1:
            mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=True))
            mods.append(dict(kind="combine", cin=NUM_CHANNELS, cout=in_ch))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=False))
    mods.append(dict(kind="attn", c=in_ch))
    mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=False, down=False))
    for lvl in reversed(range(num_res)):
        for _ in range(NUM_RES_BLOCKS + 1):
            out_ch = NF * CH_MULT[lvl]
            mods.append(dict(kind="rb", cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        if all_res[lvl] in ATTN_RESOLUTIONS:
            mods.append(dict(kind="attn", c=in_ch))
        mods.append(dict(kind="gn", c=in_ch))
        mods.append(dict(kind="conv3x3", cin=in_ch, cout=NUM_CHANNELS))
        if lvl != 0:
            mods.append(dict(kind="rb", cin=in_ch, cout=in_ch, up=True, down=False))
    assert not hs_c
    return mods


def module_params(mod: dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """(sub-name, shape) of one module's tensors in ``state_dict`` order."""
    k = mod["kind"]
    if k == "fourier":
        return [("W", (mod["size"],))]
    if k == "linear":
        return [("weight", (mod["cout"], mod["cin"])), ("bias", (mod["cout"],))]
    if k == "conv3x3":
        return [("weight", (mod["cout"], mod["cin"], 3, 3)), ("bias", (mod["cout"],))]
    if k == "gn":
        return [("weight", (mod["c"],)), ("bias", (mod["c"],))]
    if k == "combine":
        return [("Conv_0.weight", (mod["cout"], mod["cin"], 1, 1)), ("Conv_0.bias", (mod["cout"],))]
    if k == "attn":
        c = mod["c"]
        out = [("GroupNorm_0.weight", (c,)), ("GroupNorm_0.bias", (c,))]
        for i in range(4):
            out += [(f"NIN_{i}.W", (c, c)), (f"NIN_{i}.b", (c,))]
        return out
    if k == "rb":
        ci, co = mod["cin"], mod["cout"]
        out = [("GroupNorm_0.weight", (ci,)), ("GroupNorm_0.bias", (ci,)),
               ("Conv_0.weight", (co, ci, 3, 3)), ("Conv_0.bias", (co,)),
               ("Dense_0.weight", (co, TEMB_DIM)), ("Dense_0.bias", (co,)),
               ("GroupNorm_1.weight", (co,)), ("GroupNorm_1.bias", (co,)),
               ("Conv_1.weight", (co, co, 3, 3)), ("Conv_1.bias", (co,))]
        if ci != co or mod["up"] or mod["down"]:
            out += [("Conv_2.weight", (co, ci, 1, 1)), ("Conv_2.bias", (co,))]
        return out
    raise ValueError(k)


def state_dict_layout() -> List[Tuple[str, Tuple[int, ...]]]:
    """All (name, shape) pairs of ``NCSNpp.state_dict()`` in the reference's order."""
    out: List[Tuple[str, Tuple[int, ...]]] = [
        ("output_layer.weight", (2, NUM_CHANNELS, 1, 1)), ("output_layer.bias", (2,))]
    for i, mod in enumerate(module_list()):
        for sub, shape in module_params(mod):
            out.append((f"all_modules.{i}.{sub}", shape))
    return out


def num_params() -> int:
    n = 0
    for _, shape in state_dict_layout():
        k = 1
        for s in shape:
            k *= s
        n += k
    return n
