"""Integer scan at the bench shape (32 clips x 3751 x 64, interleaved, ReLU): state-split kernel under different ring depths."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, bench
    from sparsernns_b200 import fxputils, synth
    from sparsernns_b200.fxpmodel import FxpRegressionModel
    dev = torch.device("cuda:0")
    params, bs, cal = bench.load_fixture(0.5)
    xs = bench.synth_inputs(torch, 2, 32, 3751, 0, dev)
    md = fxputils.load_modeldict(params, cal)
    q = fxputils.create_fxp_qconfig(md, "w8a16", shared_exp=True)
    fx = FxpRegressionModel(md, q, 3)
    for _ in range(3): fx.forward_float(xs[0])
    print(json.dumps(bench.kernel_breakdown(lambda: fx.forward_float(xs[0]))))
else:
    envs = [{"S5B200_SCAN_I32_SPLIT": "0"}, {"S5B200_SCAN_I32_STAGES": "4"}, {"S5B200_SCAN_I32_STAGES": "20"}]
    envs += [{"S5B200_FXP_RELU_IN_GLU": "0"}]
    for env in envs:
        r = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        print(env, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
