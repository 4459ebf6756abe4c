cd $GRAFT_REPO_ROOT
echo "== 12-warp, G=1: skip none / PV mma (128) / all math (16)"
for sk in 0 128 16; do NO_BASE=1 GROUPS_N=1 LIBS=tools/probe/libs/p12.so WXB_DEC_SKIP=$sk DELAYS=0 PROF_LINES=1 bash tools/ab_groups.sh; done
echo "== 6-warp, G=2 anti-phase: skip none / PV mma / all math"
for sk in 0 128 16; do NO_BASE=1 GROUPS_N=2 LIBS=tools/probe/libs/p6.so WXB_DEC_SKIP=$sk DELAYS=240000 PROF_LINES=2 bash tools/ab_groups.sh; done
