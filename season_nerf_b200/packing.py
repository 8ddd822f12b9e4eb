"""Eval-mode program builder for the fused tcgen05 render kernel (csrc/fused_eval.cu).

The kernel is a table-driven interpreter: per tile of 128 sample points it walks a static list of MMA steps
(one weight tile x one 128x64 activation chunk each) and a static list of epilogue steps (drain an accumulator
region, bias + sin, write the next layer's activation chunks / the raw head outputs).  This module
  * folds BatchNorm (running stats) and omega_0 into the weights:  u = x.(a*30*W)^T + (a*30*b + c)
    (misc.py:188-189 with norm in eval mode; G_NeRF.py:43-50),
  * packs the bf16 weight tiles in the exact 128-byte-swizzled shared-memory image the UMMA descriptors expect,
    in consumption order, so the producer warp streams them with plain 1-D bulk copies,
  * allocates activation-chunk slots / TMEM regions and emits the barrier protocol flags,
  * checks the schedule for hazards, and provides a numpy interpreter of the tables (CPU validation of packing
    and schedule; test infrastructure for tests/test_packing.py, never a product fallback).

Layer graph (eval):  enc(63) -> fc1..fc4 -> fc5([h|enc]) -> fc6..fc9 -> X_Encode(256)
   heads: pos = [fc10Sigma; fc10Col] (4) ; solar: fc_solar_1([X_Encode|enc_sun]) -> _2 -> _3 -> fc_solar_4 (1)
   adjust: adjust_layer_1..3 -> adjust_col (12).          T_NeRF_net_v2.py:75-105, G_NeRF.py:74-133
"""
import numpy as np
import torch as t

N_SLOTS = 11            # 16 KB activation chunks resident in shared memory
N_REGIONS = 4           # TMEM accumulator regions of 128 columns
SLOT_ENC = 10           # position encoding (tile start .. fc5), then reused for enc(sun)
CHUNK = 64
NB = 128                # accumulator N-block (columns per TMEM region)

F_ACC, F_WAIT_CHUNK, F_WAIT_EMPTY, F_COMMIT = 1, 2, 4, 8
K_ENC_POS, K_ENC_SUN, K_SINE, K_HEAD = 0, 1, 2, 3
OUT_POS, OUT_VIS, OUT_ADJ = 0, 1, 2

MMA_DT = np.dtype([("w_off16", "<u4"), ("w_bytes16", "<u2"), ("a_slot", "u1"), ("n_div8", "u1"), ("d_col", "<u2"),
                   ("flags", "u1"), ("regions", "u1"), ("ksteps", "u1"), ("pad", "u1", (3,))])
EPI_DT = np.dtype([("kind", "u1"), ("region", "u1"), ("also_region", "u1"), ("ncols16", "u1"), ("d_col", "<u2"),
                   ("dst0", "u1"), ("dst1", "u1"), ("bias_off", "<u4"), ("out_id", "u1"), ("out_cols", "u1"),
                   ("pad", "u1", (2,))])
assert MMA_DT.itemsize == 16 and EPI_DT.itemsize == 16
HEADER_DT = np.dtype([("magic", "<u4"), ("n_mma", "<u4"), ("n_epi", "<u4"), ("mma_off", "<u4"), ("epi_off", "<u4"),
                      ("bias_off", "<u4"), ("w_off", "<u4"), ("total", "<u4")])
MAGIC = 0x534E4231


def fold_layer(sd, name, omega=30.0, eps=1e-5):
    """-> (W' [out,in] f32, b' [out] f32) with eval-mode BatchNorm and omega folded in."""
    W = sd[name + ".linear.weight"].detach().float().cpu()
    b = sd[name + ".linear.bias"].detach().float().cpu()
    if (name + ".norm.weight") in sd:
        inv = 1.0 / t.sqrt(sd[name + ".norm.running_var"].detach().float().cpu() + eps)
        a = sd[name + ".norm.weight"].detach().float().cpu() * inv
        c = sd[name + ".norm.bias"].detach().float().cpu() - sd[name + ".norm.running_mean"].detach().float().cpu() * a
    else:
        a, c = t.ones_like(b), t.zeros_like(b)
    return (a * omega).unsqueeze(1) * W, a * omega * b + c


def _swizzle_tile(Wt):
    """[rows, 64] float -> bytes of the bf16 SW128 K-major tile image (row r at r*128 B, 16-byte unit j at j^(r%8))."""
    rows = Wt.shape[0]
    bf = Wt.to(t.bfloat16).contiguous().view(t.int16).numpy().reshape(rows, 8, 8)
    out = np.zeros((rows, 8, 8), dtype=np.int16)
    r = np.arange(rows)
    for j in range(8):
        out[r, j ^ (r % 8)] = bf[r, j]
    return out.reshape(-1).view(np.uint8)


class _Layer:
    def __init__(self, name, W, b, in_slots, n_out, kind, out_id=None):
        self.name, self.W, self.b, self.in_slots, self.n_out, self.kind, self.out_id = name, W, b, in_slots, n_out, kind, out_id


def build_program(sd, sigma_only=False):
    """sd: state_dict of a T_NeRF(512, 4).  Returns (blob uint8 ndarray, info dict)."""
    g = "G_NeRF_net."
    lw = sd[g + "fc2.linear.weight"].shape[0]
    n_classes = sd["get_class_layer.weight"].shape[0]
    if lw != 512 or n_classes != 4:
        raise ValueError("the fused render kernel is specialised for layer_width=512, n_classes=4")

    def pad_cols(W, k):
        return t.cat([W, t.zeros(W.shape[0], k - W.shape[1])], 1) if W.shape[1] < k else W

    mma, epi, wblob, bias = [], [], [], []
    state = {"w_off": 0, "free": [s for s in range(N_SLOTS) if s != SLOT_ENC], "reg_used": [False] * N_REGIONS}

    def add_bias(v):
        off = sum(len(x) for x in bias)
        bias.append(np.asarray(v, dtype=np.float32))
        return off

    def emit_weight(Wtile):
        by = _swizzle_tile(Wtile)
        off = state["w_off"]
        wblob.append(by)
        state["w_off"] += len(by)
        assert off % 16 == 0 and len(by) % 16 == 0
        return off // 16, len(by) // 16

    def mma_step(Wtile, a_slot, d_col, n, accumulate, wait_chunk, wait_empty_region, commit_region):
        off16, by16 = emit_weight(Wtile)
        flags = (F_ACC if accumulate else 0) | (F_WAIT_CHUNK if wait_chunk else 0) | \
                (F_WAIT_EMPTY if wait_empty_region is not None else 0) | (F_COMMIT if commit_region is not None else 0)
        regs = ((wait_empty_region or 0) & 15) | (((commit_region or 0) & 15) << 4)
        mma.append((off16, by16, a_slot, n // 8, d_col, flags, regs, 4, (0, 0, 0)))

    first_read = set()     # (slot) whose current version has not been waited for yet

    def dense(W, b, in_slots, kind, out_id=None, out_cols=0, last_consumer=True, keep_inputs=()):
        """Schedule one layer.  W [n_out, 64*len(in_slots)] (already folded/padded), inputs in `in_slots`.
        Returns the list of output slots (kind == sine)."""
        n_out = W.shape[0]
        nkc = len(in_slots)
        if kind == K_HEAD:
            r = N_REGIONS - 1                      # heads use the last region
            Wp = t.cat([W, t.zeros(16 - n_out, W.shape[1])], 0)
            for kc in range(nkc):
                mma_step(Wp[:, kc * 64:(kc + 1) * 64], in_slots[kc], r * NB, 16, kc > 0, in_slots[kc] in first_read,
                         r if kc == 0 else None, r if kc == nkc - 1 else None)
                first_read.discard(in_slots[kc])
            state["reg_used"][r] = True
            epi.append((K_HEAD, r, 0xFF, 1, r * NB, 0, 0, add_bias(np.concatenate([b.numpy(), np.zeros(16 - n_out)])),
                        out_id, out_cols, (0, 0)))
            return []
        nnb = n_out // NB
        # output slots: N-block 0 drains early (while the other blocks still run) into 2 FREE slots; the later
        # blocks drain after the whole layer has retired and may overwrite this layer's own inputs.
        reusable = [s for s in in_slots if s not in keep_inputs and s != SLOT_ENC] if last_consumer else []
        early = [state["free"].pop(0), state["free"].pop(0)]
        late_pool = reusable + state["free"]
        need = 2 * (nnb - 1)
        assert len(late_pool) >= need, "out of activation slots"
        late = late_pool[:need]
        for s in late:
            if s in state["free"]:
                state["free"].remove(s)
        out_slots = early + late
        boff = add_bias(b.numpy())
        # triangular order: as soon as chunk pair p of the INPUT is ready, every N-block whose accumulator region has
        # been drained can consume it; N-block j's region is drained exactly when input pair j is ready (previous
        # layer's drain order), see DESIGN.md "fused render kernel".
        done = [[False] * nkc for _ in range(nnb)]
        order = []
        npairs = (nkc + 1) // 2
        for p in range(npairs):
            kcs = [k for k in (2 * p, 2 * p + 1) if k < nkc]
            avail_nb = min(p + 1, nnb) if p < npairs - 1 else nnb
            for j in range(avail_nb):
                for k in range(0, kcs[-1] + 1):
                    if not done[j][k]:
                        done[j][k] = True
                        order.append((j, k))
        assert all(all(d) for d in done)
        started = [False] * nnb
        remaining = [nkc] * nnb
        for (j, k) in order:
            slot = in_slots[k]
            remaining[j] -= 1
            mma_step(W[j * NB:(j + 1) * NB, k * 64:(k + 1) * 64], slot, j * NB, NB, started[j], slot in first_read,
                     j if not started[j] else None, j if remaining[j] == 0 else None)
            first_read.discard(slot)
            started[j] = True
        last_committed = order[-1][0]
        for j in range(nnb):
            state["reg_used"][j] = True
            also = 0xFF if (j == 0 or j == last_committed) else last_committed
            epi.append((K_SINE, j, also, NB // 16, j * NB, out_slots[2 * j], out_slots[2 * j + 1], boff + j * NB, 0, 0, (0, 0)))
            first_read.update(out_slots[2 * j:2 * j + 2])
        # inputs that were not overwritten and are dead return to the free list
        if last_consumer:
            for s in in_slots:
                if s not in out_slots and s not in keep_inputs and s != SLOT_ENC and s not in state["free"]:
                    state["free"].append(s)
        return out_slots

    # ---- tile prologue: position encoding into SLOT_ENC ----
    epi.append((K_ENC_POS, 0xFF, 0xFF, 0, 0, SLOT_ENC, 0, 0, 0, 0, (0, 0)))
    first_read.add(SLOT_ENC)
    W1, b1 = fold_layer(sd, g + "fc1")
    h = dense(pad_cols(W1, 64), b1, [SLOT_ENC], K_SINE, last_consumer=False)
    for name in ("fc2", "fc3", "fc4"):
        W, b = fold_layer(sd, g + name)
        h = dense(W, b, h, K_SINE)
    W5, b5 = fold_layer(sd, g + "fc5")
    W5 = pad_cols(W5, 576)
    h = dense(t.cat([W5[:, 512:576], W5[:, :512]], 1), b5, [SLOT_ENC] + h, K_SINE)       # enc chunk first: ready since tile start
    for name in ("fc6", "fc7", "fc8", "fc9"):
        W, b = fold_layer(sd, g + name)
        h = dense(W, b, h, K_SINE)
    xenc = h                                                     # 4 slots, 256 columns
    Wsig, bsig = sd[g + "fc10Sigma.weight"].float().cpu(), sd[g + "fc10Sigma.bias"].float().cpu()
    if sigma_only:
        dense(Wsig, bsig, xenc, K_HEAD, OUT_POS, 1)
    else:
        Wcol, bcol = sd[g + "fc10Col.weight"].float().cpu(), sd[g + "fc10Col.bias"].float().cpu()
        dense(t.cat([Wsig, Wcol], 0), t.cat([bsig, bcol], 0), xenc, K_HEAD, OUT_POS, 4, last_consumer=False)
        # solar branch: enc(sun) takes over the encoding slot (dead after fc5)
        epi.append((K_ENC_SUN, 0xFF, 0xFF, 0, 0, SLOT_ENC, 0, 0, 0, 0, (0, 0)))
        first_read.add(SLOT_ENC)
        Ws1, bs1 = fold_layer(sd, g + "fc_solar_1")
        Ws1 = pad_cols(Ws1, 320)
        s = dense(t.cat([Ws1[:, 256:320], Ws1[:, :256]], 1), bs1, [SLOT_ENC] + xenc, K_SINE, last_consumer=False)
        for name in ("fc_solar_2", "fc_solar_3"):
            W, b = fold_layer(sd, g + name)
            s = dense(W, b, s, K_SINE)
        dense(sd[g + "fc_solar_4.weight"].float().cpu(), sd[g + "fc_solar_4.bias"].float().cpu(), s, K_HEAD, OUT_VIS, 1)
        for x in s:
            if x not in state["free"]:
                state["free"].append(x)
        Wa, ba = fold_layer(sd, "adjust_layer_1")
        a = dense(Wa, ba, xenc, K_SINE)
        for name in ("adjust_layer_2", "adjust_layer_3"):
            W, b = fold_layer(sd, name)
            a = dense(W, b, a, K_SINE)
        dense(sd["adjust_col.weight"].float().cpu(), sd["adjust_col.bias"].float().cpu(), a, K_HEAD, OUT_ADJ, 12)

    mma_arr = np.array(mma, dtype=MMA_DT)
    epi_arr = np.array(epi, dtype=EPI_DT)
    bias_arr = np.concatenate(bias).astype(np.float32)
    w_arr = np.concatenate(wblob)
    check_schedule(mma_arr, epi_arr)

    def al(x, a=128):
        return (x + a - 1) // a * a
    hdr = np.zeros(1, dtype=HEADER_DT)
    off = al(HEADER_DT.itemsize)
    hdr["magic"], hdr["n_mma"], hdr["n_epi"] = MAGIC, len(mma_arr), len(epi_arr)
    hdr["mma_off"] = off
    off = al(off + mma_arr.nbytes)
    hdr["epi_off"] = off
    off = al(off + epi_arr.nbytes)
    hdr["bias_off"] = off
    off = al(off + bias_arr.nbytes, 1024)
    hdr["w_off"] = off
    off = al(off + w_arr.nbytes)
    hdr["total"] = off
    blob = np.zeros(off, dtype=np.uint8)
    blob[:HEADER_DT.itemsize] = hdr.view(np.uint8)
    for o, arr in ((int(hdr["mma_off"][0]), mma_arr), (int(hdr["epi_off"][0]), epi_arr), (int(hdr["bias_off"][0]), bias_arr),
                   (int(hdr["w_off"][0]), w_arr)):
        blob[o:o + arr.nbytes] = arr.view(np.uint8).reshape(-1)
    info = {"n_mma": len(mma_arr), "n_epi": len(epi_arr), "weight_bytes": int(w_arr.nbytes), "mma": mma_arr, "epi": epi_arr,
            "bias": bias_arr, "weights": w_arr}
    return blob, info


def check_schedule(mma, epi):
    """Static hazard check of the barrier protocol the kernel implements:
       (1) an epilogue step may overwrite slot s only after it has waited on an accumulator commit issued after the
           last MMA that reads the slot's previous contents;
       (2) the first MMA reading a freshly written slot carries F_WAIT_CHUNK;
       (3) an MMA that restarts a used TMEM region (accumulate=0) carries F_WAIT_EMPTY for that region;
       (4) every region commit is consumed by exactly one flipping epilogue wait, in order."""
    # map commits: for each region, the list of MMA indices that commit it, in order
    commits = {r: [] for r in range(N_REGIONS)}
    for i, m in enumerate(mma):
        if m["flags"] & F_COMMIT:
            commits[m["regions"] >> 4].append(i)
    consumed = {r: 0 for r in range(N_REGIONS)}
    observed = -1                           # MMA index up to which completion has been observed by the epilogue
    writes = []                             # (epi index, slot, observed-at-write)
    for e in epi:
        if e["kind"] in (K_SINE, K_HEAD):
            r = int(e["region"])
            assert consumed[r] < len(commits[r]), "epilogue waits on a commit that never happens"
            observed = max(observed, commits[r][consumed[r]])
            consumed[r] += 1
            if e["also_region"] != 0xFF:
                ar = int(e["also_region"])
                assert ar != r and consumed[ar] < len(commits[ar])
                observed = max(observed, commits[ar][consumed[ar]])
        if e["kind"] == K_SINE:
            writes.append((int(e["dst0"]), observed))
            writes.append((int(e["dst1"]), observed))
        elif e["kind"] in (K_ENC_POS, K_ENC_SUN):
            writes.append((int(e["dst0"]), observed))
    for r in range(N_REGIONS):
        assert consumed[r] == len(commits[r]), "unconsumed accumulator commit"
    # replay MMA order against the write order per slot
    per_slot_writes = {}
    for s, obs in writes:
        per_slot_writes.setdefault(s, []).append(obs)
    version = {s: 0 for s in range(N_SLOTS)}
    last_read = {s: -1 for s in range(N_SLOTS)}
    reg_started = [False] * N_REGIONS
    for i, m in enumerate(mma):
        s = int(m["a_slot"])
        if m["flags"] & F_WAIT_CHUNK:
            # a new version of slot s: its write must have observed completion of every read of the old version
            obs = per_slot_writes[s][version[s]]
            assert obs >= last_read[s], "slot %d overwritten before MMA %d retired" % (s, last_read[s])
            version[s] += 1
        else:
            assert version[s] > 0, "MMA reads a slot that was never written"
        last_read[s] = i
        r_wait, r_commit = m["regions"] & 15, m["regions"] >> 4
        if not (m["flags"] & F_ACC):
            region = int(m["d_col"]) // NB
            assert m["flags"] & F_WAIT_EMPTY and r_wait == region, "region (re)started without waiting for its drain"
        else:
            assert not (m["flags"] & F_WAIT_EMPTY)
    for s, lst in per_slot_writes.items():
        assert version[s] == len(lst), "slot %d: %d writes but %d waited versions" % (s, len(lst), version[s])
    return True


# ---------------------------------------------------------------------------------------------------------
# numpy interpreter of the tables (validation of packing + schedule on the CPU; not a product path)
def _unswizzle_tile(raw, rows):
    x = raw.view(np.int16).reshape(rows, 8, 8)
    out = np.zeros_like(x)
    r = np.arange(rows)
    for j in range(8):
        out[r, j] = x[r, j ^ (r % 8)]
    return t.from_numpy(out.reshape(rows, 64).copy()).view(t.bfloat16).float()


def interpret(info, enc_pos, enc_sun):
    """enc_pos [n,64], enc_sun [n,64] (already padded) float32 -> {out_id: raw head outputs}.  Sequential emulation
    of the two roles: before each epilogue step every MMA up to the commit(s) it waits on is executed; an MMA that
    would need a slot version not yet written means the real kernel would deadlock (asserted).  bf16 activation
    storage, fp32 accumulation, like the kernel."""
    n = enc_pos.shape[0]
    mma, epi, bias, w = info["mma"], info["epi"], info["bias"], info["weights"]
    commits = {r: [] for r in range(N_REGIONS)}
    for i, m in enumerate(mma):
        if m["flags"] & F_COMMIT:
            commits[int(m["regions"]) >> 4].append(i)
    consumed = {r: 0 for r in range(N_REGIONS)}
    slots = [None] * N_SLOTS
    written = [0] * N_SLOTS
    waited = [0] * N_SLOTS
    regions = t.zeros(n, N_REGIONS * NB)
    outs = {}
    bf = lambda x: x.to(t.bfloat16).float()
    state = {"mi": 0}
    drained = [0] * N_REGIONS
    empty_waits = [0] * N_REGIONS

    def run_to(target):
        while state["mi"] <= target:
            m = mma[state["mi"]]
            sl = int(m["a_slot"])
            if m["flags"] & F_WAIT_CHUNK:
                assert written[sl] > waited[sl], "deadlock: MMA %d waits for slot %d that is written later" % (state["mi"], sl)
                waited[sl] += 1
            if m["flags"] & F_WAIT_EMPTY:
                rw = int(m["regions"]) & 15
                assert drained[rw] >= empty_waits[rw], "deadlock: MMA %d waits for a drain of region %d that comes later" % (state["mi"], rw)
                empty_waits[rw] += 1
            rows = int(m["n_div8"]) * 8
            raw = w[int(m["w_off16"]) * 16:(int(m["w_off16"]) + int(m["w_bytes16"])) * 16]
            contrib = slots[sl] @ _unswizzle_tile(raw, rows).T
            c0 = int(m["d_col"])
            if m["flags"] & F_ACC:
                regions[:, c0:c0 + rows] += contrib
            else:
                regions[:, c0:c0 + rows] = contrib
            state["mi"] += 1

    for e in epi:
        k = int(e["kind"])
        if k == K_ENC_POS:
            slots[int(e["dst0"])] = bf(enc_pos)
            written[int(e["dst0"])] += 1
            continue
        if k == K_ENC_SUN:
            slots[int(e["dst0"])] = bf(enc_sun)
            written[int(e["dst0"])] += 1
            continue
        r = int(e["region"])
        target = commits[r][consumed[r]]
        consumed[r] += 1
        if e["also_region"] != 0xFF:
            ar = int(e["also_region"])
            target = max(target, commits[ar][consumed[ar]])
        run_to(target)
        c0, nc = int(e["d_col"]), int(e["ncols16"]) * 16
        acc = regions[:, c0:c0 + nc] + t.from_numpy(bias[int(e["bias_off"]):int(e["bias_off"]) + nc].copy())
        drained[r] += 1
        if k == K_SINE:
            y = bf(t.sin(acc))
            for d, sl in ((0, int(e["dst0"])), (1, int(e["dst1"]))):
                slots[sl] = y[:, 64 * d:64 * d + 64].clone()
                written[sl] += 1
        else:
            outs[int(e["out_id"])] = acc[:, :int(e["out_cols"])].clone()
    return outs
