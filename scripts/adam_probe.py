import torch as t
t.manual_seed(0)
dev = "cuda"
def run(fused, graph):
    t.manual_seed(1)
    ps = [t.nn.Parameter(t.randn(n, device=dev)) for n in (7, 1000, 33)]
    lr = t.tensor(1e-3, device=dev)
    opt = t.optim.Adam(ps, lr=lr, capturable=True, fused=fused)
    sched = t.optim.lr_scheduler.OneCycleLR(opt, max_lr=1e-3, total_steps=20, base_momentum=0.85, max_momentum=0.95, cycle_momentum=False)
    gs = [[t.randn_like(p) for p in ps] for _ in range(6)]
    v0 = ps[0]._version
    def step(i):
        for p, g in zip(ps, gs[i]):
            p.grad = g.clone() if p.grad is None else p.grad.copy_(g)
        opt.step()
    step(0); sched.step()
    if graph:
        static_g = [t.zeros_like(p) for p in ps]
        for p, g in zip(ps, static_g): p.grad = g
        s = t.cuda.Stream(); s.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(s):
            for sg, g in zip(static_g, gs[1]): sg.copy_(g)
            opt.step()
        t.cuda.current_stream().wait_stream(s); sched.step()
        G = t.cuda.CUDAGraph()
        for sg, g in zip(static_g, gs[2]): sg.copy_(g)
        with t.cuda.graph(G):
            opt.step()
        sched.step()
        for i in (3, 4, 5):
            for sg, g in zip(static_g, gs[i]): sg.copy_(g)
            G.replay(); sched.step()
    else:
        for i in range(1, 6):
            step(i); sched.step()
    t.cuda.synchronize()
    return [p.detach().clone() for p in ps], ps[0]._version - v0, [float(opt.state[p]["step"]) for p in ps], float(opt.param_groups[0]["lr"])
ref = run(False, False)
for fused in (False, True):
    for graph in (False, True):
        r = run(fused, graph)
        print("fused", fused, "graph", graph, "max diff vs foreach-eager", max(float((a - b).abs().max()) for a, b in zip(r[0], ref[0])), "version bumps", r[1], "steps", r[2], "lr", r[3])
