"""CPU: network descriptions (tables -> tensor graph) and the header parser."""
import os

import numpy as np
import pytest

from tf2_b200 import nets
from tf2_b200.netdesc import NetDesc


def test_builtin_resnet50_graph():
    n = nets.resnet50()
    assert n.num_layers == 54 and len(n.tensors) == 55
    assert n.macs_per_image() == 4089184256          # 4.0892 GMAC
    l0 = n.layers[0]
    assert (l0.C, l0.k, l0.OH, l0.pool, l0.pool_stride, l0.pool_pad, l0.PH) == (27, 3, 112, 1, 2, 1, 56)
    # residual operands: res2a_2c adds branch1, res2b_2c adds res2a's output, ...
    assert n.layers[4].add_tensor == n.layers[1].out_tensor
    assert n.layers[7].add_tensor == n.layers[4].out_tensor
    assert n.layers[14].add_tensor == n.layers[11].out_tensor
    assert sum(1 for l in n.layers if l.add_tensor >= 0) == 16
    # stride-2 layers carry their real output size
    assert (n.layers[13].stride, n.layers[13].OH) == (2, 28) and (n.layers[11].stride, n.layers[11].OH) == (2, 28)
    assert n.layers[52].gap == 1 and n.layers[53].N == 1000 and n.layers[53].bias_en == 1
    # only the image can hold -128
    assert [l.in_may_be_m128 for l in n.layers] == [1] + [0] * 53


def test_builtin_googlenet_graph():
    n = nets.googlenet()
    assert n.num_layers == 67 and n.macs_per_image() == 1582671872
    assert sum(l.ipool for l in n.layers) == 9
    t = n.tensors[n.layers[3].out_tensor]
    assert t.C == 256 and [n.layers[i].out_ch0 for i in (3, 5, 7, 9)] == [0, 64, 192, 224]
    assert n.tensors[n.result_tensor()].C == 1000
    assert all(l.out_ch0 % 16 == 0 for l in n.layers)


@pytest.mark.parametrize("name", ["resnet50", "googlenet", "resnet50_pruned"])
def test_json_matches_reference_header(name, reference_dir):
    hdr = os.path.join(reference_dir, "Runtime_Engine/cnn/host/inc", name + ".h")
    a = NetDesc.from_header(hdr, name).to_json()
    b = nets.load(name).to_json()
    assert a == b


def test_header_parser_macros():
    from tf2_b200.header_tables import parse_header
    t = parse_header("""
      #define NUM_LAYER 2
      #define FOO (CEIL(27, C_VECTOR) * 2)   // comment
      CONSTANT int kA[NUM_LAYER] = { CEIL(114, W_VECTOR), /* x */ NEXT_POWER_OF_2(48) };
      CONSTANT bool kB[NUM_LAYER] = { 1, 0 };
      CONSTANT int kAMax = 17;
    """)
    assert t["FOO"] == 4 and t["kA"] == [17, 64] and t["kB"] == [1, 0] and t["kAMax"] == 17


def test_chain_builder_flags():
    n = nets.chain((8, 6, 6), [dict(N=8, k=1, relu=0), dict(N=8, k=3, pad=1), dict(ipool=1, src=0)])
    assert [l.in_may_be_m128 for l in n.layers] == [1, 1, 1]   # layer 0 has no ReLU -> its output may be -128
    n = nets.chain((8, 6, 6), [dict(N=8, k=1), dict(N=8, k=3, pad=1)])
    assert [l.in_may_be_m128 for l in n.layers] == [1, 0]
