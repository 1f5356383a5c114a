"""Long-form synthesis: the window/chain logic of tools/longform_synthesis.py:258-478 (BASELINE config 5).

A stream of L frames is cut into 150-frame windows with a 15-frame overlap
(`chunk_starts = [0] + range(135, L, 135)`, last window zero-padded, :263-287); annotations are re-based
to window time (:348-381); window c is sampled with `use_prev_latent=True, prev_latent=` the output latent
of window c-1 (:389-404), so SAMPLING is a serial chain ("replicas only") while retrieval and the
DDIM inversions of all windows do not depend on the chain.  `LongformSynthesizer.run` therefore
  1. prepares every window (retrieval, exemplar fetch/encode) and inverts ALL windows' exemplars in one
     batched 50-step loop (`batch_inversions=True`; the reference inverts per window, per exemplar, B=1),
  2. runs the guided sampling chain window by window,
  3. cross-fades consecutive windows linearly over the overlap: rotations in 6D space, expressions and
     translation directly (:431-478).
On several GPUs (torch.distributed initialised) step 1 is SHARDED: rank r prepares and inverts
the windows shard_range(n_windows, r, world), turns the inverted latents into per-level insertion targets
(MotionDiffusion.insertion_targets: deterministic, no RNG), and ONE all-gather (parallel.gather_clips: ~6 MB per
window over NVLink) hands every window's targets and pre-projected conditions to the chain rank, which runs step
2 (B = 1, CUDA-graph replay of the evaluation chain) and step 3.  The other ranks are free for other streams
("replicas only" for the chain).
BERT / wav2vec feature extraction per window is out of scope: the caller supplies per-window features.
"""
import torch

from . import config as CFG


def chunk_starts(n_frames, window=CFG.MAX_SEQ_LEN, overlap=15):
    """[0] + range(window - overlap, n_frames, window - overlap)  (longform_synthesis.py:263)."""
    return [0] + list(range(window - overlap, n_frames, window - overlap))


def rebase_annotations(discourse, prominence, gesture_labels, t0, t1):
    """Keep the annotations lying inside [t0, t1] seconds and shift them to window time (:348-381)."""
    disc = [(d[0], d[1], d[2], d[3], d[4] - t0, d[5] - t0, d[6] - t0, d[7] - t0)
            for d in discourse if d[4] >= t0 and d[5] <= t1]
    prom = [(p[0], p[1] - t0, p[2] - t0, p[3]) for p in prominence if p[1] >= t0 and p[2] <= t1]
    gest = [{"start": g["start"] - t0, "end": g["end"] - t0, "name": g["name"], "word": g["word"]}
            for g in gesture_labels if g["start"] >= t0 and g["end"] <= t1]
    return disc, prom, gest


# ---- rotation helpers for the 6D cross-fade (standard formulas; torch, any device) -----------------------
def axis_angle_to_matrix(aa):
    """Rodrigues: [..., 3] -> [..., 3, 3]."""
    angle = aa.norm(dim=-1, keepdim=True)
    small = angle < 1e-6
    axis = aa / torch.where(small, torch.ones_like(angle), angle)
    x, y, z = axis.unbind(-1)
    zero = torch.zeros_like(x)
    K = torch.stack([zero, -z, y, z, zero, -x, -y, x, zero], -1).reshape(aa.shape[:-1] + (3, 3))
    s, c = torch.sin(angle)[..., None], torch.cos(angle)[..., None]
    eye = torch.eye(3, dtype=aa.dtype, device=aa.device).expand(K.shape)
    # K @ K = a a^T - (a . a) I for the skew matrix of a; written out: a batched 3x3 matmul over every joint of
    # every frame is a 260 us cuBLAS call where three elementwise kernels take 15
    KK = axis[..., :, None] * axis[..., None, :] - (axis * axis).sum(-1)[..., None, None] * eye
    R = eye + s * K + (1 - c) * KK
    return torch.where(small[..., None], eye, R)


def matrix_to_rotation_6d(m):
    return m[..., :2, :].clone().reshape(m.shape[:-2] + (6,))


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = torch.nn.functional.normalize(a1, dim=-1)
    b2 = torch.nn.functional.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def matrix_to_axis_angle(R):
    """[..., 3, 3] -> [..., 3] (angle in [0, pi]); quaternion route for stability near 0 and pi."""
    m00, m11, m22 = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    qw = 0.5 * torch.sqrt(torch.clamp(1 + m00 + m11 + m22, min=0))
    qx = 0.5 * torch.sqrt(torch.clamp(1 + m00 - m11 - m22, min=0)) * torch.sign(R[..., 2, 1] - R[..., 1, 2] + 1e-30)
    qy = 0.5 * torch.sqrt(torch.clamp(1 - m00 + m11 - m22, min=0)) * torch.sign(R[..., 0, 2] - R[..., 2, 0] + 1e-30)
    qz = 0.5 * torch.sqrt(torch.clamp(1 - m00 - m11 + m22, min=0)) * torch.sign(R[..., 1, 0] - R[..., 0, 1] + 1e-30)
    v = torch.stack([qx, qy, qz], -1)
    n = v.norm(dim=-1, keepdim=True)
    angle = 2 * torch.atan2(n, qw[..., None])
    return torch.where(n < 1e-8, torch.zeros_like(v), v / n.clamp_min(1e-30) * angle)


def crossfade_rotations(prev_tail, new_head):
    """Linear blend over the overlap in 6D space (:449-471): [B, F, J*3] axis-angle in and out."""
    B, F, D = new_head.shape
    w = torch.linspace(0, 1, F, device=new_head.device, dtype=new_head.dtype).view(1, F, 1)
    to6 = lambda a: matrix_to_rotation_6d(axis_angle_to_matrix(a.reshape(B, F, D // 3, 3))).reshape(B, F, D // 3 * 6)
    blended = to6(prev_tail) * (1 - w) + to6(new_head) * w
    return matrix_to_axis_angle(rotation_6d_to_matrix(blended.reshape(B, F, D // 3, 6))).reshape(B, F, D)


def crossfade_linear(prev_tail, new_head):
    F = new_head.shape[1]
    w = torch.linspace(0, 1, F, device=new_head.device, dtype=new_head.dtype).view(1, F, 1)
    return prev_tail * (1 - w) + new_head * w


class LongformSynthesizer:
    """Drives a MotionDiffusion over one stream.  `window_fn(cidx, f0, f1)` returns the batch dict of window
    cidx (frames [f0, f1), schema of beatx_collate_fn with B=1; per-window audio/text features included)."""

    ROT_KEYS = ("pred_upper", "pred_lower", "pred_hands", "pred_facepose")
    LIN_KEYS = ("pred_exps", "pred_transl")

    def __init__(self, arch, window=CFG.MAX_SEQ_LEN, overlap=15, fps=CFG.MOTION_FPS):
        self.arch, self.window, self.overlap, self.fps = arch, window, overlap, fps

    @staticmethod
    def _window_names(batch, cidx):
        """Every window must retrieve its own exemplars: RetrievalDatabase.retrieve caches by sample_name, and
        the reference renames each chunk ('/0' -> '/{cidx}', tools/longform_synthesis.py:347).  Applied here so a
        window_fn that returns the stream's name for every window cannot silently reuse window 0's exemplars."""
        names = batch.get("sample_name")
        if names is None:
            return batch
        single = isinstance(names, str)
        out = []
        for nm in ([names] if single else list(names)):
            head, sep, tail = str(nm).rpartition("/")
            out.append(f"{head}/{cidx}" if sep and tail.isdigit() else f"{nm}/{cidx}")
        batch["sample_name"] = out[0] if single else out
        return batch

    # ---- window payload: what the chain rank needs from the rank that prepared the window ------------------
    PAYLOAD = ("inv_per_t", "start_rows", "start_mask", "xf_text", "xf_audio", "xf_spk", "motion_mask")

    def pack_window(self, gb):
        """One prepared + inverted window (B = 1) as a flat fp32 vector: insertion targets, pre-projected
        conditions and the motion mask.  Returns (vector, shapes)."""
        arch = self.arch
        arch.encode_clip_conditions(gb)
        start_rows, start_mask, inv_per_t = arch.insertion_targets(gb)
        xf = gb.model_kwargs["xf_out"]
        parts = dict(inv_per_t=inv_per_t, start_rows=start_rows, start_mask=start_mask.float(),
                     xf_text=xf["xf_text"], xf_audio=xf["xf_audio"], xf_spk=xf["xf_spk"],
                     motion_mask=gb.model_kwargs["motion_mask"])
        shapes = {k: tuple(parts[k].shape) for k in self.PAYLOAD}
        return torch.cat([parts[k].float().reshape(-1) for k in self.PAYLOAD]), shapes

    def unpack_window(self, vec, shapes, template):
        """The GuidedBatch of a window prepared elsewhere: `template` (any locally prepared window) supplies the
        flags; tensors come from the payload."""
        from .architecture import GuidedBatch
        out, off = {}, 0
        for k in self.PAYLOAD:
            n = 1
            for d in shapes[k]:
                n *= d
            out[k] = vec[off:off + n].reshape(shapes[k]).clone()     # own storage: the kernels need 16-byte aligned rows
            off += n
        gb = GuidedBatch()
        for f in ("use_outpaint", "use_inversion", "inversion_start_time", "use_guidance", "guidance_iters", "guidance_lr",
                  "shape", "device"):
            setattr(gb, f, getattr(template, f))
        gb.use_prev, gb.prev_latent, gb.extra, gb.jobs, gb.results = True, None, {}, (), {}
        qm = template.model_kwargs["query_mask"]
        gb.model_kwargs = dict(xf_out={k: out[k].contiguous() for k in ("xf_text", "xf_audio", "xf_spk")}, re_dict=None,
                               sample_idx=None, query_mask=qm, motion_mask=out["motion_mask"].contiguous())
        gb.targets = (out["start_rows"].contiguous(), out["start_mask"] > 0.5, out["inv_per_t"].contiguous())
        return gb

    def merge_windows(self, vecs, shapes, template):
        """ONE GuidedBatch of S clips from the payloads of S windows (the same window index of S independent
        streams): the chain of several streams then runs as one batch per window -- at B = 1 an evaluation is
        latency-bound (~100 dependent kernels), so S streams cost about the time of one."""
        gbs = [self.unpack_window(v, shapes, template) for v in vecs]
        if len(gbs) == 1:
            return gbs[0]
        gb = gbs[0]
        S = len(gbs)
        gb.shape = (S,) + tuple(template.shape[1:])
        xf = {k: torch.cat([g.model_kwargs["xf_out"][k] for g in gbs], 0) for k in ("xf_text", "xf_audio", "xf_spk")}
        qm = {c: m[:1].expand(S, -1).contiguous() for c, m in template.model_kwargs["query_mask"].items()}
        gb.model_kwargs = dict(xf_out=xf, re_dict=None, sample_idx=None, query_mask=qm,
                               motion_mask=torch.cat([g.model_kwargs["motion_mask"] for g in gbs], 0))
        gb.targets = (torch.cat([g.targets[0] for g in gbs], 0), torch.cat([g.targets[1] for g in gbs], 0),
                      torch.cat([g.targets[2] for g in gbs], 1))
        return gb

    def run_streams(self, n_frames, window_fns, inference_kwargs):
        """S independent streams of the same length on ONE GPU: every stream's windows are prepared, all their
        exemplars inverted in one batched loop, and the S prev-latent chains advance together, one batch of S clips
        per window index.  Returns the result dict with a leading stream dimension (pred_* [S, frames, ...],
        latents [n_windows * S, 43, 512] window-major).  Per stream the arithmetic equals run(batch_inversions=True);
        the start noise of a window is drawn for the S streams at once."""
        starts = chunk_starts(n_frames, self.window, self.overlap)
        arch = self.arch
        per_stream = []
        db = getattr(arch.model, "database", None)
        for si, fn in enumerate(window_fns):
            # _window_names gives window c of EVERY stream the name ".../c": the retrieval cache (keyed by name) must
            # not serve one stream's exemplars to another
            for cache in ((db.test_indexes, db.test_dbounds, db.test_qbounds) if db is not None else ()):
                cache.clear()
            gbs = []
            for cidx, f0 in enumerate(starts):
                batch = self._window_names(dict(fn(cidx, f0, f0 + self.window)), cidx)
                batch["inference_kwargs"] = dict(inference_kwargs, use_prev_latent=True, prev_latent=None)
                gbs.append(arch.prepare(**batch))
            per_stream.append(gbs)
        arch.invert_many([gb for gbs in per_stream for gb in gbs])
        packed = [[self.pack_window(gb) for gb in gbs] for gbs in per_stream]
        shapes, template = packed[0][0][1], per_stream[0][0]
        merged = [self.merge_windows([packed[si][c][0] for si in range(len(per_stream))], shapes, template)
                  for c in range(len(starts))]
        return self._chain(merged, starts)

    def _chain(self, windows, starts):
        """Steps 2 and 3: the serial sampling chain over prepared windows (exemplars inverted or targets shipped)
        and the cross-fade of consecutive windows."""
        arch = self.arch
        outs, prev, latents = None, None, []
        for item in windows:
            item.use_prev, item.prev_latent = True, arch.mask_prev_latent(prev)
            res = arch.finish(item, arch.run_prepared(item))
            prev = res["prev_latentout"]
            latents.append(prev)
            outs = self._append(outs, {k: res[k] for k in self.ROT_KEYS + self.LIN_KEYS})
        outs["latents"] = torch.cat(latents, 0)
        outs["window_starts"] = starts
        return outs

    def _append(self, outs, cur):
        if outs is None:
            return cur
        ov = self.overlap
        for k in cur:
            fade = crossfade_rotations if k in self.ROT_KEYS else crossfade_linear
            head = fade(outs[k][:, -ov:], cur[k][:, :ov])
            outs[k] = torch.cat([outs[k][:, :-ov], head, cur[k][:, ov:]], dim=1)
        return outs

    def run_sharded(self, n_frames, window_fn, inference_kwargs, group=None, chain_rank=0, timings=None):
        """Multi-GPU form of run(batch_inversions=True) (module docstring): windows' retrieval + inversions sharded
        over the ranks, one all-gather of the window payloads, the chain on `chain_rank`.  Returns the result dict
        on the chain rank and None elsewhere.  With one process it equals run(batch_inversions=True).
        `timings` (a dict) receives host-clock seconds of the phases, each closed by a device synchronise."""
        import time
        from .parallel import _world, gather_clips, shard_range

        def stamp(name, t0):
            if timings is not None:
                torch.cuda.synchronize()
                timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
            return time.perf_counter()
        t = time.perf_counter()
        rank, world = _world(group)
        starts = chunk_starts(n_frames, self.window, self.overlap)
        n_win = len(starts)
        lo, hi = shard_range(n_win, rank, world)
        arch = self.arch
        mine = []
        for cidx in range(lo, hi):
            batch = self._window_names(dict(window_fn(cidx, starts[cidx], starts[cidx] + self.window)), cidx)
            batch["inference_kwargs"] = dict(inference_kwargs, use_prev_latent=True, prev_latent=None)
            mine.append(arch.prepare(**batch))
        if not mine:
            raise RuntimeError("run_sharded: more ranks than windows; use fewer ranks for this stream")
        t = stamp("prepare", t)
        arch.invert_many(mine)                      # this rank's exemplars, one batched reverse loop
        t = stamp("invert", t)
        if world == 1:
            out = self._chain(mine, starts)
            stamp("chain", t)
            return out
        packed = [self.pack_window(gb) for gb in mine]
        shapes = packed[0][1]
        local = torch.stack([v for v, _ in packed], 0)
        full = gather_clips(local, n_win, group)    # [n_win, payload] on every rank, window order
        t = stamp("exchange", t)
        if rank != chain_rank:
            return None
        windows = [mine[c - lo] if lo <= c < hi else self.unpack_window(full[c], shapes, mine[0]) for c in range(n_win)]
        for gb in mine:                             # shipped and local windows take the same code path
            gb.targets, gb.inv = arch.insertion_targets(gb), None
        out = self._chain(windows, starts)
        stamp("chain", t)
        return out

    def run(self, n_frames, window_fn, inference_kwargs, batch_inversions=True):
        """batch_inversions=True prepares every window (codec encodes, retrieval) and inverts all exemplars in
        one batched reverse loop BEFORE the first start_noise is drawn; the reference interleaves encode /
        start_noise / sampling window by window, so same-seed runs of the two modes draw different noise (each
        mode is reproducible on its own; batch_inversions=False keeps the reference's order)."""
        starts = chunk_starts(n_frames, self.window, self.overlap)
        arch = self.arch
        prepared = []
        for cidx, f0 in enumerate(starts):
            batch = self._window_names(dict(window_fn(cidx, f0, f0 + self.window)), cidx)
            ik = dict(inference_kwargs, use_prev_latent=True, prev_latent=None)
            batch["inference_kwargs"] = ik
            if not batch_inversions:
                prepared.append(batch)
                continue
            gb = arch.prepare(**batch)
            prepared.append(gb)
        if batch_inversions:
            arch.invert_many(prepared)              # one batched reverse loop for every window's exemplars
            return self._chain(prepared, starts)
        outs, prev, latents = None, None, []
        for item in prepared:                       # the reference's order: window by window
            item["inference_kwargs"]["prev_latent"] = prev
            res = arch(**item)
            prev = res["prev_latentout"]
            latents.append(prev)
            outs = self._append(outs, {k: res[k] for k in self.ROT_KEYS + self.LIN_KEYS})
        outs["latents"] = torch.cat(latents, 0)
        outs["window_starts"] = starts
        return outs
