import sys, os
sys.path.insert(0, "/root/repo")
sys.argv = ["x"]
import tools.tc_profile as P
for (R, H) in [(256, 160), (32, 240)]:
    for T in (8, 16, 32, 63, 125, 501):
        P.run(R, H, T)
