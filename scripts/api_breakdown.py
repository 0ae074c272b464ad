import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import coverage_control as cc
from coverage_control.algorithms import DecentralizedCVT, ClairvoyantCVT
from bench import synth_inputs, make_idf, CNN
p = cc.Parameters()
bnd, pos = synth_inputs(0)
idf = make_idf(cc, p, bnd, device=0)
idf.GenerateMap()
for Ctl in (DecentralizedCVT, ClairvoyantCVT):
    env = cc.CoverageSystem(p, idf, cc.PointVector(pos))
    ctrl = Ctl(p, env)
    acc = [0.0]*5
    n = 300
    for it in range(n+20):
        if it == 20: acc = [0.0]*5
        t0 = time.perf_counter(); ctrl.ComputeActions()
        t1 = time.perf_counter(); a = ctrl.GetActions()
        t2 = time.perf_counter(); env.StepActions(a)
        t3 = time.perf_counter(); env.GetObjectiveValue()
        t4 = time.perf_counter(); cc.CoverageEnvUtils.get_maps(env, p, CNN, True)
        t5 = time.perf_counter()
        for i,(x,y) in enumerate(((t0,t1),(t1,t2),(t2,t3),(t3,t4),(t4,t5))): acc[i] += y-x
    print(Ctl.__name__, {k: round(v/n*1e6,1) for k,v in zip(("ComputeActions","GetActions","StepActions","GetObjectiveValue","get_maps"),acc)}, "total", round(sum(acc)/n*1e6,1))
