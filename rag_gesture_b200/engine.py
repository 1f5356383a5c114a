"""Thin owner of one `rg_handle`: everything arithmetic happens in librg_b200.so.

PyTorch is used for device memory, streams and RNG only.  A DenoiserEngine is created from a
reference-layout state dict (keys of ReGestureTransformer without the leading "model."), holds the
packed device weights, the timestep table (K7) and serves per-batch cross-attention state (K6).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import config as CFG


class DenoiserEngine:
    def __init__(self, state_dict, latent_dim=CFG.LATENT_DIM, num_heads=CFG.NUM_HEADS,
                 ffn_dim=CFG.FF_SIZE, time_embed_dim=CFG.TIME_EMBED_DIM, num_layers=CFG.NUM_LAYERS,
                 n_chunks=CFG.N_CHUNKS, text_dim=CFG.TEXT_DIM, num_speakers=CFG.NUM_SPEAKERS,
                 precision=_lib.PREC_FP32, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("rg_b200: no CUDA device; the hot path has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_tokens = 4 * n_chunks + 3
        self.latent_dim = latent_dim
        self.num_layers = num_layers
        cfg = _lib.RgConfig(latent_dim, num_heads, ffn_dim, time_embed_dim, num_layers,
                            self.n_tokens, n_chunks, text_dim, num_speakers, precision)
        keep, names, ptrs, numels = [], [], [], []
        for k, v in state_dict.items():
            if k.startswith("gesture_rep_encoder."):
                continue
            t = v.detach().to(device="cpu", dtype=torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            numels.append(t.numel())
        n = len(names)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rg_create(C.byref(cfg), n, (C.c_char_p * n)(*names),
                                          (C.c_void_p * n)(*ptrs), (C.c_int64 * n)(*numels),
                                          C.byref(h)))
        self._h = h
        self.state_floats_per_clip = int(self.lib.rg_state_floats_per_clip(self._h))
        self.n_steps = 0

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.rg_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- schedule -------------------------------------------------------------------------------
    def set_schedule(self, timestep_map, coef):
        """coef: float32 [S,8] as documented at rg_set_schedule in include/rg_b200.h."""
        tm = np.ascontiguousarray(np.asarray(timestep_map, dtype=np.int32))
        cf = np.ascontiguousarray(np.asarray(coef, dtype=np.float32))
        assert cf.shape == (len(tm), 8)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rg_set_schedule(
                self._h, len(tm), tm.ctypes.data_as(C.POINTER(C.c_int32)),
                cf.ctypes.data_as(C.POINTER(C.c_float)), _lib.stream_ptr()))
        self.n_steps = len(tm)
        self.timestep_map = [int(t) for t in tm]

    # -- once per clip ----------------------------------------------------------------------------
    def encode_conditions(self, word=None, audio=None, speaker_ids=None):
        """-> xf_text [B,Nt,512], xf_audio [B,Na,512], xf_spk [B,Ns,512] (raggesture.py:978-987).
        Any input may be None; its output is then absent."""
        _lib.require_cuda(word, audio, speaker_ids)
        D = self.latent_dim
        word = None if word is None else word.float().contiguous()
        audio = None if audio is None else audio.float().contiguous()
        spk = None if speaker_ids is None else speaker_ids.to(torch.int64).contiguous()
        ref = next(t for t in (word, audio, spk) if t is not None)
        B, dev = ref.shape[0], ref.device
        xt = None if word is None else torch.empty(B, word.shape[1], D, device=dev)
        xa = None if audio is None else torch.empty(B, audio.shape[1], D, device=dev)
        xs = None if spk is None else torch.empty(B, spk.shape[1], D, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.rg_encode_conditions(
                self._h, _lib.ptr(word), _lib.ptr(audio), _lib.ptr(spk), B,
                0 if word is None else word.shape[1], 0 if audio is None else audio.shape[1],
                0 if spk is None else spk.shape[1], _lib.ptr(xt), _lib.ptr(xa), _lib.ptr(xs),
                _lib.stream_ptr()))
        out = {"xf_text": xt, "xf_audio": xa, "xf_spk": xs}
        return {k: v for k, v in out.items() if v is not None}

    def precompute_state(self, xf_out):
        """K6: [B, L, 3, 16, 32, 32] fp32 cross-attention state for a batch of clips."""
        xt, xa, xs = (xf_out[k].float().contiguous() for k in CFG.CONDS)
        _lib.require_cuda(xt, xa, xs)
        B = xt.shape[0]
        state = torch.empty(B, self.state_floats_per_clip, device=xt.device)
        with torch.cuda.device(xt.device):
            _lib.check(self.lib.rg_precompute_clip_state(
                self._h, _lib.ptr(xt), _lib.ptr(xa), _lib.ptr(xs), xt.shape[1], xa.shape[1],
                xs.shape[1], B, _lib.ptr(state), _lib.stream_ptr()))
        return state.view(B, self.num_layers, 3, CFG.NUM_HEADS, 32, 32)

    def set_lanes(self, lanes):
        """Concurrent clip-range chains per denoise call: 0 = automatic, 1..4 fixed (rg_set_lanes)."""
        _lib.check(self.lib.rg_set_lanes(self._h, int(lanes)))

    def set_graphs(self, on):
        """CUDA-graph replay of the evaluation chain (rg_set_graphs); on by default."""
        _lib.check(self.lib.rg_set_graphs(self._h, int(bool(on))))

    # -- per step -----------------------------------------------------------------------------------
    def denoise(self, x, src_mask, query_mask, state, step_idx=-1, tau=0, out=None):
        """x0 = model(x, t) for B clips sharing one timestep.  query_mask: [3,B,T] tensor or None."""
        _lib.require_cuda(x, src_mask, state)
        x = x.float().contiguous()
        B = x.shape[0]
        assert x.shape[1] == self.n_tokens and x.shape[2] == self.latent_dim, x.shape
        assert state.shape[0] == B, "state batch does not match x"
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(self.lib.rg_denoise(
                self._h, _lib.ptr(x), B, int(step_idx), int(tau), _lib.ptr(src_mask),
                _lib.ptr(query_mask), _lib.ptr(state), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def denoise_groups(self, x, src_mask, query_mask, state, groups, out=None):
        """One evaluation of a batch made of consecutive clip ranges at DIFFERENT schedule levels:
        groups = [(n_clips, step_idx), ...] (rg_denoise_groups).  Per clip identical to denoise()."""
        _lib.require_cuda(x, src_mask, state)
        assert x.is_contiguous() and x.dtype == torch.float32
        B = x.shape[0]
        assert state.shape[0] == B and sum(n for n, _ in groups) == B, "groups / state do not match x"
        if out is None:
            out = torch.empty_like(x)
        n = len(groups)
        clips = (C.c_int32 * n)(*[int(g[0]) for g in groups])
        steps = (C.c_int32 * n)(*[int(g[1]) for g in groups])
        with torch.cuda.device(x.device):
            _lib.check(self.lib.rg_denoise_groups(self._h, _lib.ptr(x), B, n, clips, steps, _lib.ptr(src_mask),
                                                  _lib.ptr(query_mask), _lib.ptr(state), _lib.ptr(out),
                                                  _lib.stream_ptr()))
        return out

    def run_levels(self, S, xj, B, E, src_mask, query_mask, state, in_seq0=None, inv_list=None, noise=None,
                   guidance_iters=None, guidance_lr=0.1, run_dead_guidance=False, samples_out=None):
        """rg_run_levels: S levels of guided/plain sampling for xj[:B] and of DDIM inversion for xj[B:B+E] in one
        C call (one kernel chain per level for both ranges).  xj is updated in place."""
        _lib.require_cuda(xj, src_mask, state)
        for t in (xj, in_seq0, inv_list, noise, samples_out):
            assert t is None or (t.is_contiguous() and t.dtype == torch.float32)
        assert xj.shape[0] == B + E == state.shape[0] == src_mask.shape[0]
        assert inv_list is None or tuple(inv_list.shape) == (S, B) + tuple(xj.shape[1:])
        assert noise is None or tuple(noise.shape) == (S, B) + tuple(xj.shape[1:])
        assert E == 0 or tuple(samples_out.shape) == (S, E) + tuple(xj.shape[1:])
        gi = None
        if guidance_iters is not None:
            assert len(guidance_iters) >= S
            gi = (C.c_int32 * len(guidance_iters))(*[int(g) for g in guidance_iters])
        x0 = torch.empty_like(xj)
        with torch.cuda.device(xj.device):
            _lib.check(self.lib.rg_run_levels(self._h, int(S), _lib.ptr(xj), int(B), int(E), _lib.ptr(in_seq0),
                                              _lib.ptr(inv_list), _lib.ptr(noise), gi, float(guidance_lr),
                                              int(bool(run_dead_guidance)), _lib.ptr(src_mask), _lib.ptr(query_mask),
                                              _lib.ptr(state), _lib.ptr(samples_out), _lib.ptr(x0), _lib.stream_ptr()))
        return xj

    @staticmethod
    def _f32(*tensors, like=None):
        """The kernels read raw fp32: refuse other dtypes / layouts / sizes instead of reading out of bounds."""
        for t in tensors:
            if t is None:
                continue
            _lib.require_cuda(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"rg_b200: expected a contiguous float32 CUDA tensor, got {t.dtype}, contiguous={t.is_contiguous()}")
            if like is not None and t.numel() != like.numel():
                raise ValueError(f"rg_b200: tensor of {t.numel()} elements where {like.numel()} are expected")

    def ddim_update(self, x, x0, step_idx, direction, out=None):
        if out is None:
            out = torch.empty_like(x)
        self._f32(x, x0, out, like=x)
        with torch.cuda.device(x.device):
            _lib.check(self.lib.rg_ddim_update(self._h, _lib.ptr(x), _lib.ptr(x0), int(step_idx),
                                               int(direction), _lib.ptr(out), x.numel(),
                                               _lib.stream_ptr()))
        return out

    def blend_in_seq(self, x, in_seq, noise, step_idx, out=None):
        if out is None:
            out = torch.empty_like(x)
        self._f32(x, in_seq, noise, out, like=x)
        rows = x.numel() // self.latent_dim
        with torch.cuda.device(x.device):
            _lib.check(self.lib.rg_blend_in_seq(self._h, _lib.ptr(x), _lib.ptr(in_seq),
                                                _lib.ptr(noise), int(step_idx), _lib.ptr(out),
                                                rows, _lib.stream_ptr()))
        return out

    def mix_branches(self, out2, coefs, joint_scale, out=None):
        """rg_mix_branches: out2 [2B,T,D] (text rows, then "none" rows), coefs [B,4], joint_scale [T] -> [B,T,D]."""
        self._f32(out2, coefs, joint_scale)
        B = out2.shape[0] // 2
        assert coefs.shape == (B, 4) and joint_scale.numel() == self.n_tokens
        if out is None:
            out = torch.empty((B,) + tuple(out2.shape[1:]), device=out2.device)
        with torch.cuda.device(out2.device):
            _lib.check(self.lib.rg_mix_branches(self._h, _lib.ptr(out2), B, _lib.ptr(coefs), _lib.ptr(joint_scale),
                                                _lib.ptr(out), _lib.stream_ptr()))
        return out

    def guidance_steps(self, x, in_seq, iters, lr, numel=None):
        self._f32(x, in_seq, like=x)
        rows = x.numel() // self.latent_dim
        with torch.cuda.device(x.device):
            _lib.check(self.lib.rg_guidance_steps(self._h, _lib.ptr(x), _lib.ptr(in_seq), rows,
                                                  int(iters), float(lr),
                                                  int(numel if numel is not None else x.numel()),
                                                  _lib.stream_ptr()))
        return x


def launch_count():
    return int(_lib.load().rg_launch_count())
