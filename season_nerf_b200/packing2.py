"""Eval-mode program builder for the CTA-pair fused render kernel (csrc/fused_eval2.cu).

Two CTAs of a cluster render a tile of 256 sample points (128 rows each) with tcgen05.mma.cta_group::2: every
MMA step multiplies one [256 x 64] activation chunk (128 rows in each CTA's shared memory) with one weight tile
[n x 64] of which each CTA stages HALF (n/2 rows, TMA, 128B swizzle applied by the TMA unit).  Accumulators live in
two 256-column TMEM regions per CTA.  Per tile the kernel walks
  * a static list of MMA steps   (weight rows, activation slot, TMEM column, barrier flags), and
  * a static list of epilogue steps (encode inputs / drain one region: bias + sin -> bf16 -> next layer's chunks /
    heads -> global memory),
both generated here together with the folded weights:  u = x.(a*30*W)^T + (a*30*b + c)  (eval-mode BatchNorm and
omega_0 folded in; misc.py:188-189, G_NeRF.py:43-50).

Shared-memory plan per CTA: 9 activation slots of 16 KB (slot 8 = positional encoding, later enc(sun)) + 5 weight
stages of 16 KB.  512-wide layers run in place:  with input halves A|B the MMA order is (blk0,A) (blk1,A) (blk0,B)
-> commit0, (blk1,B) -> commit1; the drain of block 0 overwrites the A slots (dead once commit0 has fired), the drain
of block 1 the B slots, and the next layer starts on the A chunks while block 1 is still being drained.

Layer graph (eval):  enc(63) -> fc1..fc4 -> fc5([h|enc]) -> fc6..fc9 -> X_Encode(256)
   heads: pos = [fc10Sigma; fc10Col] (4) ; solar: fc_solar_1([X_Encode|enc_sun]) -> _2 -> _3 -> fc_solar_4 (1)
   adjust: adjust_layer_1..3 -> adjust_col (12).          T_NeRF_net_v2.py:75-105, G_NeRF.py:74-133

`interpret` / `check_schedule` are CPU validation of packing and schedule (tests/test_packing.py), never a product
fallback.
"""
import numpy as np
import torch as t


N_SLOTS = 9
N_REGIONS = 2
SLOT_ENC = 8
CHUNK = 64
NB = 256                # accumulator N-block = TMEM region width
HEAD_N = 16

F_ACC, F_WAIT_CHUNK, F_WAIT_EMPTY, F_COMMIT = 1, 2, 4, 8
K_ENC_POS, K_ENC_SUN, K_SINE, K_HEAD = 0, 1, 2, 3
OUT_POS, OUT_VIS, OUT_ADJ = 0, 1, 2

MMA_DT = np.dtype([("w_row", "<u4"), ("n", "<u2"), ("a_slot", "u1"), ("flags", "u1"), ("d_col", "<u2"), ("regions", "u1"),
                   ("pad", "u1", (5,))])
EPI_DT = np.dtype([("kind", "u1"), ("region", "u1"), ("also_region", "u1"), ("nchunks", "u1"), ("d_col", "<u2"),
                   ("out_id", "u1"), ("out_cols", "u1"), ("bias_off", "<u4"), ("dst", "u1", (4,))])
assert MMA_DT.itemsize == 16 and EPI_DT.itemsize == 16
HEADER_DT = np.dtype([("magic", "<u4"), ("n_mma", "<u4"), ("n_epi", "<u4"), ("mma_off", "<u4"), ("epi_off", "<u4"),
                      ("bias_off", "<u4"), ("w_off", "<u4"), ("w_rows", "<u4"), ("total", "<u4")])
MAGIC = 0x534E4232


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


def build_program(sd, sigma_only=False):
    """sd: state_dict of a T_NeRF(512, 4).  Returns (blob uint8 ndarray, info dict)."""
    g = "G_NeRF_net."
    lw = sd[g + "fc2.linear.weight"].shape[0]
    n_classes = sd["get_class_layer.weight"].shape[0]
    if lw != 512 or n_classes != 4:
        raise ValueError("the fused render kernel is specialised for layer_width=512, n_classes=4")

    def pad_cols(W, k):
        return t.cat([W, t.zeros(W.shape[0], k - W.shape[1])], 1) if W.shape[1] < k else W

    mma, epi, wrows, bias = [], [], [], []
    state = {"rows": 0, "last_use": [-1] * N_REGIONS}
    first_read = set()          # slots whose current version no MMA has waited for yet

    def add_bias(v):
        off = sum(len(x) for x in bias)
        bias.append(np.asarray(v, dtype=np.float32))
        return off

    def emit_weight(Wtile):
        """[n, 64] float -> appended to the weight matrix (bf16 rows of 128 bytes); returns the first row index.
        Tiles start on 8-row boundaries (one swizzle atom) so that both TMA box heights (128 / 8) stay aligned."""
        assert Wtile.shape[1] == CHUNK
        r0 = state["rows"]
        wrows.append(Wtile.to(t.bfloat16).contiguous())
        state["rows"] += Wtile.shape[0]
        return r0

    def mma_step(Wtile, a_slot, d_col, accumulate, wait_empty_region, commit_region):
        n = Wtile.shape[0]
        wait_chunk = a_slot in first_read
        first_read.discard(a_slot)
        flags = (F_ACC if accumulate else 0) | (F_WAIT_CHUNK if wait_chunk else 0) | \
                (F_WAIT_EMPTY if wait_empty_region is not None else 0) | (F_COMMIT if commit_region is not None else 0)
        regs = ((wait_empty_region or 0) & 15) | (((commit_region or 0) & 15) << 4)
        mma.append((emit_weight(Wtile), n, a_slot, flags, d_col, regs, (0,) * 5))
        if commit_region is not None:
            state["last_use"][commit_region] = len(mma) - 1

    def pick_region():
        return int(np.argmin(state["last_use"]))

    def head(W, b, in_slots, out_id, out_cols):
        n_out = W.shape[0]
        r = pick_region()
        Wp = t.cat([W, t.zeros(HEAD_N - n_out, W.shape[1])], 0)
        nkc = len(in_slots)
        for kc in range(nkc):
            mma_step(Wp[:, kc * 64:(kc + 1) * 64], in_slots[kc], r * NB, kc > 0, r if kc == 0 else None,
                     r if kc == nkc - 1 else None)
        epi.append((K_HEAD, r, 0xFF, 0, r * NB, out_id, out_cols,
                    add_bias(np.concatenate([b.numpy(), np.zeros(HEAD_N - n_out)])), (0, 0, 0, 0)))

    def sine(W, b, in_slots, out_slots):
        """One SIREN layer.  W [n_out, 64*len(in_slots)] folded/padded; out_slots: n_out/64 slots in column order."""
        n_out = W.shape[0]
        nkc = len(in_slots)
        nnb = n_out // NB
        assert n_out % NB == 0 and len(out_slots) == n_out // CHUNK
        boff = add_bias(b.numpy())
        if nnb == 1:
            r = pick_region()
            order = [(0, k) for k in range(nkc)]
            regions = [r]
        else:
            assert nnb == 2
            regions = [0, 1]
            ha = (nkc + 1) // 2
            A, B = list(range(ha)), list(range(ha, nkc))
            order = [(0, k) for k in A] + [(1, k) for k in A] + [(0, k) for k in B] + [(1, k) for k in B]
        started = [False] * nnb
        remaining = [nkc] * nnb
        commit_order = []
        for (j, k) in order:
            remaining[j] -= 1
            r = regions[j]
            mma_step(W[j * NB:(j + 1) * NB, k * 64:(k + 1) * 64], in_slots[k], r * NB, started[j],
                     r if not started[j] else None, r if remaining[j] == 0 else None)
            started[j] = True
            if remaining[j] == 0:
                commit_order.append(j)
        for j in commit_order:
            dst = tuple(out_slots[4 * j:4 * j + 4])
            epi.append((K_SINE, regions[j], 0xFF, 4, regions[j] * NB, 0, 0, boff + j * NB, dst))
            first_read.update(dst)

    H = list(range(8))
    X = [0, 1, 2, 3]
    T4 = [4, 5, 6, 7]
    # ---- tile prologue: position encoding into SLOT_ENC ----
    epi.append((K_ENC_POS, 0xFF, 0xFF, 0, 0, 0, 0, 0, (SLOT_ENC, 0, 0, 0)))
    first_read.add(SLOT_ENC)
    W1, b1 = fold_layer(sd, g + "fc1")
    sine(pad_cols(W1, 64), b1, [SLOT_ENC], H)
    for name in ("fc2", "fc3", "fc4"):
        W, b = fold_layer(sd, g + name)
        sine(W, b, H, H)
    W5, b5 = fold_layer(sd, g + "fc5")
    W5 = pad_cols(W5, 576)
    sine(t.cat([W5[:, 512:576], W5[:, :512]], 1), b5, [SLOT_ENC] + H, H)          # enc chunk first: ready since tile start
    for name in ("fc6", "fc7", "fc8"):
        W, b = fold_layer(sd, g + name)
        sine(W, b, H, H)
    W9, b9 = fold_layer(sd, g + "fc9")
    sine(W9, b9, H, X)                                                            # X_Encode: slots 0..3
    Wsig, bsig = sd[g + "fc10Sigma.weight"].float().cpu(), sd[g + "fc10Sigma.bias"].float().cpu()
    if sigma_only:
        head(Wsig, bsig, X, OUT_POS, 1)
    else:
        Wcol, bcol = sd[g + "fc10Col.weight"].float().cpu(), sd[g + "fc10Col.bias"].float().cpu()
        head(t.cat([Wsig, Wcol], 0), t.cat([bsig, bcol], 0), X, OUT_POS, 4)
        # solar branch: enc(sun) takes over the encoding slot (dead after fc5)
        epi.append((K_ENC_SUN, 0xFF, 0xFF, 0, 0, 0, 0, 0, (SLOT_ENC, 0, 0, 0)))
        first_read.add(SLOT_ENC)
        Ws1, bs1 = fold_layer(sd, g + "fc_solar_1")
        Ws1 = pad_cols(Ws1, 320)
        sine(t.cat([Ws1[:, 256:320], Ws1[:, :256]], 1), bs1, [SLOT_ENC] + X, T4)
        for name in ("fc_solar_2", "fc_solar_3"):
            W, b = fold_layer(sd, g + name)
            sine(W, b, T4, T4)
        head(sd[g + "fc_solar_4.weight"].float().cpu(), sd[g + "fc_solar_4.bias"].float().cpu(), T4, OUT_VIS, 1)
        # seasonal adjust branch: block 0 of adjust_layer_1 drains into the (dead) solar slots, block 1 over X_Encode
        A1 = T4 + X
        Wa, ba = fold_layer(sd, "adjust_layer_1")
        sine(Wa, ba, X, A1)
        for name in ("adjust_layer_2", "adjust_layer_3"):
            W, b = fold_layer(sd, name)
            sine(W, b, A1, A1)
        head(sd["adjust_col.weight"].float().cpu(), sd["adjust_col.bias"].float().cpu(), A1, OUT_ADJ, 12)

    mma_arr = np.array(mma, dtype=MMA_DT)
    epi_arr = np.array(epi, dtype=EPI_DT)
    bias_arr = np.concatenate(bias).astype(np.float32)
    w_mat = t.cat(wrows, 0)                                      # [rows, 64] bf16
    w_arr = w_mat.view(t.int16).numpy().reshape(-1).view(np.uint8)
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
    hdr["w_rows"] = w_mat.shape[0]
    off = al(off + w_arr.nbytes)
    hdr["total"] = off
    blob = np.zeros(off, dtype=np.uint8)
    blob[:HEADER_DT.itemsize] = hdr.view(np.uint8)
    for o, arr in ((int(hdr["mma_off"][0]), mma_arr), (int(hdr["epi_off"][0]), epi_arr), (int(hdr["bias_off"][0]), bias_arr),
                   (int(hdr["w_off"][0]), w_arr)):
        blob[o:o + arr.nbytes] = arr.view(np.uint8).reshape(-1)
    info = {"n_mma": len(mma_arr), "n_epi": len(epi_arr), "weight_bytes": int(w_arr.nbytes), "w_rows": int(w_mat.shape[0]),
            "mma": mma_arr, "epi": epi_arr, "bias": bias_arr, "weights": w_mat.float()}
    return blob, info


def check_schedule(mma, epi):
    """Static hazard check of the barrier protocol the kernel implements:
       (1) an epilogue step may overwrite slot s only after it has waited on an accumulator commit issued after the
           last MMA that reads the slot's previous contents;
       (2) the first MMA reading a freshly written slot carries F_WAIT_CHUNK;
       (3) an MMA that restarts a TMEM region (accumulate=0) carries F_WAIT_EMPTY for that region;
       (4) every region commit is consumed by exactly one epilogue wait, in order;
       (5) a region is never restarted before the epilogue step draining its previous contents (deadlock-free order
           is verified by `interpret`)."""
    commits = {r: [] for r in range(N_REGIONS)}
    for i, m in enumerate(mma):
        if m["flags"] & F_COMMIT:
            commits[int(m["regions"]) >> 4].append(i)
    consumed = {r: 0 for r in range(N_REGIONS)}
    observed = -1                           # MMA index up to which completion has been observed by the epilogue
    writes = []                             # (slot, observed-at-write)
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
            for d in e["dst"][:int(e["nchunks"])]:
                writes.append((int(d), observed))
        elif e["kind"] in (K_ENC_POS, K_ENC_SUN):
            writes.append((int(e["dst"][0]), observed))
    for r in range(N_REGIONS):
        assert consumed[r] == len(commits[r]), "unconsumed accumulator commit"
    per_slot_writes = {}
    for s, obs in writes:
        per_slot_writes.setdefault(s, []).append(obs)
    version = {s: 0 for s in range(N_SLOTS)}
    last_read = {s: -1 for s in range(N_SLOTS)}
    for i, m in enumerate(mma):
        s = int(m["a_slot"])
        assert s < N_SLOTS
        if m["flags"] & F_WAIT_CHUNK:
            obs = per_slot_writes[s][version[s]]
            assert obs >= last_read[s], "slot %d overwritten before MMA %d retired" % (s, last_read[s])
            version[s] += 1
        else:
            assert version[s] > 0, "MMA reads a slot that was never written"
        last_read[s] = i
        r_wait = int(m["regions"]) & 15
        if not (m["flags"] & F_ACC):
            region = int(m["d_col"]) // NB
            assert m["flags"] & F_WAIT_EMPTY and r_wait == region, "region (re)started without waiting for its drain"
        else:
            assert not (m["flags"] & F_WAIT_EMPTY)
        assert int(m["d_col"]) + int(m["n"]) <= N_REGIONS * NB and int(m["n"]) in (HEAD_N, NB)
    for s, lst in per_slot_writes.items():
        assert version[s] == len(lst), "slot %d: %d writes but %d waited versions" % (s, len(lst), version[s])
    # across tiles: the first writes of the next tile (encoding, fc1 drains) must not race the last reads of this
    # tile: every slot's final read has to be observed by the last epilogue step of the tile
    assert observed >= max(last_read.values()), "tile tail: an MMA may still read a slot when the next tile starts"
    return True


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
            rows = int(m["n"])
            Wt = w[int(m["w_row"]):int(m["w_row"]) + rows]
            contrib = slots[sl] @ Wt.T
            c0 = int(m["d_col"])
            if m["flags"] & F_ACC:
                regions[:, c0:c0 + rows] += contrib
            else:
                regions[:, c0:c0 + rows] = contrib
            state["mi"] += 1

    for e in epi:
        k = int(e["kind"])
        if k == K_ENC_POS:
            slots[int(e["dst"][0])] = bf(enc_pos)
            written[int(e["dst"][0])] += 1
            continue
        if k == K_ENC_SUN:
            slots[int(e["dst"][0])] = bf(enc_sun)
            written[int(e["dst"][0])] += 1
            continue
        r = int(e["region"])
        target = commits[r][consumed[r]]
        consumed[r] += 1
        if e["also_region"] != 0xFF:
            ar = int(e["also_region"])
            target = max(target, commits[ar][consumed[ar]])
        run_to(target)
        c0 = int(e["d_col"])
        drained[r] += 1
        if k == K_SINE:
            nc = 64 * int(e["nchunks"])
            acc = regions[:, c0:c0 + nc] + t.from_numpy(bias[int(e["bias_off"]):int(e["bias_off"]) + nc].copy())
            y = bf(t.sin(acc))
            for d in range(int(e["nchunks"])):
                sl = int(e["dst"][d])
                slots[sl] = y[:, 64 * d:64 * d + 64].clone()
                written[sl] += 1
        else:
            acc = regions[:, c0:c0 + HEAD_N] + t.from_numpy(bias[int(e["bias_off"]):int(e["bias_off"]) + HEAD_N].copy())
            outs[int(e["out_id"])] = acc[:, :int(e["out_cols"])].clone()
    return outs
