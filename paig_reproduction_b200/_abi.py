"""ctypes mirror of include/paig_b200.h (struct layouts and argument types).  No library is loaded here."""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 2
CELL_IDS = {"spring": 0, "bouncing": 1, "gravity": 2}


class Task(C.Structure):
    _fields_ = [("cell", C.c_int32), ("n_objs", C.c_int32), ("H", C.c_int32), ("seq_len", C.c_int32),
                ("input_steps", C.c_int32), ("pred_steps", C.c_int32), ("alt_vel", C.c_int32),
                ("deep_unet", C.c_int32), ("alpha", C.c_float), ("batch_global", C.c_int32),
                ("gravity_A", C.c_float), ("flags", C.c_int32)]


FLAG_INFERENCE = 1


class WB(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("content_l1", WB), ("content_l2", WB), ("background_l1", WB), ("background_l2", WB),
                ("template_l1", WB), ("template_l2", WB), ("conv", WB * 18), ("enc_l1", WB), ("enc_l2", WB),
                ("enc_l3", WB), ("vel", WB * 3), ("dt", C.c_void_p), ("phys0", C.c_void_p), ("phys1", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("output_seq", C.c_void_p), ("recons_out", C.c_void_p), ("enc_pos", C.c_void_p),
                ("pos_vel_seq", C.c_void_p), ("enc_masks", C.c_void_p), ("masked_objs", C.c_void_p),
                ("templates", C.c_void_p), ("losses", C.c_void_p)]


# state_dict key prefix -> (field, index or None).  `unet` is "shallow_unet" or "unet".
def param_slots(unet: str, n_convs: int, alt_vel: bool, cell: str):
    slots = {
        "var_net_content.l1": ("content_l1", None), "var_net_content.l2": ("content_l2", None),
        "var_net_background.l1": ("background_l1", None), "var_net_background.l2": ("background_l2", None),
        "var_net_template.l1": ("template_l1", None), "var_net_template.l2": ("template_l2", None),
        "encoder.l1": ("enc_l1", None), "encoder.l2": ("enc_l2", None), "encoder.l3": ("enc_l3", None),
    }
    for i in range(n_convs):
        slots["encoder.%s.c%d" % (unet, i + 1)] = ("conv", i)
    if alt_vel:
        slots["velocity_encoder.init_vel_linear"] = ("vel", 0)
    else:
        for j, k in enumerate((0, 2, 4)):
            slots["velocity_encoder.init_vel_mlp.%d" % k] = ("vel", j)
    scalars = {"rollout_cell.dt": "dt"}
    if cell == "spring":
        scalars.update({"rollout_cell.k": "phys0", "rollout_cell.equil": "phys1"})
    elif cell == "gravity":
        scalars.update({"rollout_cell.g": "phys0", "rollout_cell.m": "phys1"})
    return slots, scalars


def fill_params(struct: Params, ptr_of, names, unet: str, n_convs: int, alt_vel: bool, cell: str):
    """ptr_of(name) -> integer address or None; names: iterable of state_dict keys to bind."""
    slots, scalars = param_slots(unet, n_convs, alt_vel, cell)
    for name in names:
        addr = ptr_of(name)
        if name in scalars:
            setattr(struct, scalars[name], addr)
            continue
        prefix, _, leaf = name.rpartition(".")
        if prefix not in slots or leaf not in ("weight", "bias"):
            continue
        field, idx = slots[prefix]
        wb = getattr(struct, field) if idx is None else getattr(struct, field)[idx]
        setattr(wb, "w" if leaf == "weight" else "b", addr)
    return struct


def declare(lib):
    """Set restype/argtypes on a loaded library (the real one or the test shim build)."""
    vp, i, l = C.c_void_p, C.c_int, C.c_long
    PT, PP, PO = C.POINTER(Task), C.POINTER(Params), C.POINTER(Outputs)
    sig = {
        "paig_abi_version": (C.c_int, []),
        "paig_last_error": (C.c_char_p, []),
        "paig_workspace_bytes": (C.c_size_t, [PT, i]),
        "paig_step_forward": (i, [PT, PP, vp, i, PO, vp, vp]),
        "paig_step_backward": (i, [PT, PP, PP, vp, i, vp, vp, vp, vp, vp, vp]),
        "paig_step_fused": (i, [PT, PP, PP, vp, i, PO, vp, vp]),
        "paig_step_fused_host": (i, [PT, PP, PP, vp, i, vp, vp, vp]),
        "paig_set_early_grad_event": (None, [vp]),
        "paig_rollout_forward": (i, [i, i, i, i, vp, vp, vp, vp, vp]),
        "paig_rollout_backward": (i, [i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp]),
        "paig_templates_forward": (i, [PT, PP, vp, vp, vp, vp]),
        "paig_templates_backward": (i, [PT, PP, PP, vp, vp, vp, vp, vp]),
        "paig_decode_forward": (i, [PT, vp, vp, i, vp, vp, l, i, vp, vp]),
        "paig_decode_backward": (i, [PT, vp, vp, i, vp, vp, l, i, vp, vp, vp, vp, vp, vp]),
        "paig_decode_layers": (i, [PT, vp, vp, i, vp, vp, vp]),
        "paig_encoder_forward": (i, [PT, PP, vp, l, i, i, vp, vp, vp, vp, vp]),
        "paig_encoder_backward": (i, [PT, PP, PP, vp, l, i, i, vp, vp, vp]),
        "paig_velocity_forward": (i, [PT, PP, vp, i, vp, vp, vp]),
        "paig_velocity_backward": (i, [PT, PP, PP, vp, i, vp, vp, vp, vp]),
        "paig_launch_count": (C.c_long, []),
        "paig_profile_begin": (None, []),
        "paig_profile_end": (i, [C.c_char_p, C.c_size_t]),
        "paig_frame_sse_forward": (i, [vp, l, i, vp, i, i, i, vp, vp]),
        "paig_frame_sse_backward": (i, [vp, l, i, vp, i, i, i, vp, vp, vp]),
        "paig_debug_workspace_offset": (C.c_long, [PT, i, C.c_char_p, i]),
        "paig_debug_unet_conv_view": (i, [PT, i, i, C.POINTER(C.c_long * 5)]),
        "paig_conv3x3_forward": (i, [vp, vp, vp, vp, i, i, i, i, i, vp]),
        "paig_conv3x3_backward": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp]),
        "paig_optimizer_step": (i, [i, vp, vp, vp, vp, l, C.c_float, i, vp]),
        "paig_optimizer_step_f64": (i, [i, vp, vp, vp, vp, l, C.c_double, i, vp]),
        "paig_gather_batch_u8": (i, [vp, l, vp, i, vp, vp]),
        "paig_debug_gemm_tc": (i, [vp, vp, vp, i, i, i, i, vp, l, vp]),
        "paig_debug_conv3x3_tc": (i, [vp, vp, vp, vp, i, i, i, i, i, i, vp, vp]),
        "paig_stage_input_host": (i, [PT, vp, i, i, vp, vp]),
        "paig_step_fused_staged": (i, [PT, PP, PP, i, i, vp, vp, vp]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    return missing


EXPORTS = ["paig_abi_version", "paig_last_error", "paig_workspace_bytes", "paig_step_forward", "paig_step_backward",
           "paig_step_fused", "paig_step_fused_host", "paig_set_early_grad_event", "paig_rollout_forward", "paig_rollout_backward",
           "paig_templates_forward", "paig_templates_backward", "paig_decode_forward", "paig_decode_backward", "paig_decode_layers",
           "paig_encoder_forward", "paig_encoder_backward", "paig_velocity_forward", "paig_velocity_backward",
           "paig_conv3x3_forward", "paig_conv3x3_backward", "paig_debug_workspace_offset", "paig_debug_unet_conv_view",
           "paig_frame_sse_forward", "paig_frame_sse_backward", "paig_launch_count", "paig_profile_begin",
           "paig_profile_end", "paig_optimizer_step", "paig_optimizer_step_f64", "paig_gather_batch_u8", "paig_debug_gemm_tc", "paig_debug_conv3x3_tc", "paig_stage_input_host",
           "paig_step_fused_staged"]
