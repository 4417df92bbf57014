import sys, os
sys.argv=["x"]
exec(open("profiles/time_create.py").read().split("for probes in")[0])
for rep in range(3):
    os.environ["BSREG_TIMING"]="2"
    h = capi.Handle(yp, Xp, model=capi.MODEL_SAR, W=Wp, ldet=capi.LDET_CHEBYSHEV, cheb_order=50, cheb_probes=64, n_chains=64, seed=1 + rep)
    h.close()
    print("----")
