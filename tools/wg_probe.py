import sys, torch, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import backends, stage_checks as sc
be = backends.get("cuda")
import ast
for (N,Cin,Cout,S) in [ast.literal_eval(sys.argv[1])]:
    try:
        sc.check_conv3x3(be, N, Cin, Cout, S, False)
        print("ok", N,Cin,Cout,S, flush=True)
    except Exception as e:
        print("FAIL", N,Cin,Cout,S, str(e)[:200], flush=True)
        break
