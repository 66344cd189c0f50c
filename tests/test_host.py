"""CPU-side checks of the host layer: the C-ABI library loads and exports every symbol the header declares,
the drop-in module mirrors the reference's state_dict, and the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from oracle import physicsnet_oracle as po
from paig_reproduction_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER_ARGS = {  # positional order of runners/torch_run_physics.py:81-84
    "spring_color": ("spring_color", 100, 1, "spring_ode_cell", 12, 4, 6, 3.0, False, True, 32 * 32, "conv_encoder", "conv_st_decoder"),
    "bouncing_balls": ("bouncing_balls", 100, 1, "bouncing_ode_cell", 12, 4, 6, 2.0, False, True, 32 * 32, "conv_encoder", "conv_st_decoder"),
    "3bp_color": ("3bp_color", 100, 1, "gravity_ode_cell", 20, 4, 12, 5.0, False, True, 36 * 36, "conv_encoder", "conv_st_decoder"),
    "mnist_spring_color": ("mnist_spring_color", 100, 1, "spring_ode_cell", 12, 3, 7, 3.0, False, True, 64 * 64, "conv_encoder", "conv_st_decoder"),
}


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "paig_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(paig_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from paig_reproduction_b200 import build
    build.build()                                            # nvcc cross-compiles sm_100a without a GPU
    lib = _lib.load()                                        # raises if anything declared in _abi is missing
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_abi.EXPORTS) == names                     # the ctypes mirror and the header agree
    assert lib.paig_abi_version() == _abi.ABI_VERSION


def test_workspace_query_needs_no_gpu():
    lib = _lib.load()
    tk = _abi.Task(0, 2, 32, 12, 4, 6, 0, 0, 3.0, 0)
    small, big = lib.paig_workspace_bytes(ctypes.byref(tk), 1), lib.paig_workspace_bytes(ctypes.byref(tk), 100)
    assert 0 < small < big < 4 << 30
    bad = _abi.Task(0, 2, 32, 10, 4, 6, 0, 0, 3.0, 0)       # seq_len must exceed in+pr (physics_models.py:59)
    assert lib.paig_workspace_bytes(ctypes.byref(bad), 1) == 0
    assert b"invalid" in lib.paig_last_error()


@pytest.mark.parametrize("task", list(RUNNER_ARGS))
def test_state_dict_mirrors_reference(task):
    from paig_reproduction_b200.physics_models import PhysicsNet
    net = PhysicsNet(*RUNNER_ARGS[task], device="cpu")
    sd = net.state_dict()
    ref = po.param_shapes(po.TASKS[task])
    assert list(sd.keys()) == [k for k, _, _ in ref]
    for k, shape, dtype in ref:
        assert tuple(sd[k].shape) == shape and sd[k].dtype == dtype, k
    live = net.live_parameter_names()
    assert not any(k.startswith("rollout_cell.weight") or k.endswith(".dt") or k.endswith(".m") for k in live)
    dead_unet = "encoder.shallow_unet." if po.TASKS[task].H >= 40 else "encoder.unet."
    assert not any(k.startswith(dead_unet) for k in live)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout only exists in the build container")
def test_default_init_equals_reference_under_same_seed():
    """Same construction order => the drop-in starts from the reference's weights for a given torch seed."""
    from oracle.make_golden import build_reference_net, import_reference
    from paig_reproduction_b200.physics_models import PhysicsNet
    pm = import_reference()
    spec = po.TASKS["spring_color"]
    torch.manual_seed(3)
    ref = build_reference_net(pm, spec, 12, 3.0, False).state_dict()
    torch.manual_seed(3)
    ours = PhysicsNet(*RUNNER_ARGS["spring_color"], device="cpu").state_dict()
    assert list(ref.keys()) == list(ours.keys())
    for k in ref:
        assert torch.equal(ref[k], ours[k]), k


def test_no_cpu_fallback():
    from paig_reproduction_b200.physics_models import PhysicsNet
    net = PhysicsNet(*RUNNER_ARGS["spring_color"], device="cpu")
    x = torch.rand(1, 12, 3, 32, 32)
    with pytest.raises(_lib.PaigError):
        net(x)
    with pytest.raises(_lib.PaigError):
        net.train_step(x)


def test_constructor_errors_match_reference():
    from paig_reproduction_b200.physics_models import PhysicsNet
    a = list(RUNNER_ARGS["spring_color"])
    with pytest.raises(AssertionError):
        PhysicsNet(*(a[:4] + [10] + a[5:]), device="cpu")   # seq_len == in+pr
    with pytest.raises(KeyError):
        PhysicsNet(*(a[:3] + ["lstm_cell"] + a[4:]), device="cpu")
    with pytest.raises(AssertionError):
        PhysicsNet(*(["no_such_task"] + a[1:]), device="cpu")


@pytest.mark.parametrize("task", list(RUNNER_ARGS))
def test_flat_gradient_order_puts_the_unet_last(task):
    """parallel.py all-reduces the prefix of the flat gradient buffer underneath the UNet backward: that prefix must
    hold every live fp32 parameter except the UNet conv layers, and be the bulk of the bytes."""
    from paig_reproduction_b200.physics_models import PhysicsNet
    net = PhysicsNet(*RUNNER_ARGS[task], device="cpu")
    order = net.flat_order()
    params = dict(net.named_parameters())
    live = [k for k in net.live_parameter_names() if params[k].dtype == torch.float32]
    assert sorted(order) == sorted(live) and len(set(order)) == len(order)
    conv = "encoder." + ("unet." if task == "mnist_spring_color" else "shallow_unet.")
    first_conv = min(i for i, k in enumerate(order) if k.startswith(conv))
    assert all(k.startswith(conv) for k in order[first_conv:]) and not any(k.startswith(conv) for k in order[:first_conv])
    early = sum(params[k].numel() for k in order[:first_conv])
    total = sum(params[k].numel() for k in order)
    assert early / total > (0.9 if task != "mnist_spring_color" else 0.85)


@pytest.mark.parametrize("task,alt_vel", [(t, False) for t in RUNNER_ARGS] + [("spring_color", True)])
def test_parameter_table_slot_index_equals_the_field_by_field_fill(task, alt_vel):
    """PhysicsNet fills paig_params through a slot index resolved once per net (physics_models._slot_of) instead of ctypes
    attribute access per call; byte for byte it must be the table _abi.fill_params builds, for every task / cell / velocity
    encoder variant, and the live-gradient allocation must hand out disjoint, correctly shaped, 16-byte aligned views."""
    from paig_reproduction_b200.physics_models import PhysicsNet
    args = list(RUNNER_ARGS[task])
    args[8] = alt_vel
    net = PhysicsNet(*args, device="cpu")
    names = [k for k, _ in net.named_parameters()]
    fake = {k: 0x7f0000000000 + 4096 * i for i, k in enumerate(names)}
    want = _abi.Params()
    _abi.fill_params(want, lambda k: fake[k], names, net._unet, net._n_convs, net.alt_vel, net.cell_kind)
    got = _abi.Params()
    arr = (ctypes.c_uint64 * (ctypes.sizeof(_abi.Params) // 8)).from_buffer(got)
    slot = net._slot_of()
    for k in names:
        if k in slot:
            arr[slot[k]] = fake[k]
    assert bytes(want) == bytes(got)
    assert len(set(slot.values())) == len(slot)
    # live parameters all have a slot; parameters without one are exactly those the step never touches
    assert all(k in slot for k in net.live_parameter_names())
    params = {k: v.data for k, v in net.named_parameters()}
    live = net.live_parameter_names()
    grads = net._alloc_grads(params, live)
    spans = []
    for k in live:
        g = grads[k]
        assert g.shape == params[k].shape and g.dtype == params[k].dtype and g.is_contiguous()
        assert g.data_ptr() % 16 == 0
        spans.append((g.data_ptr(), g.data_ptr() + g.numel() * g.element_size()))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
