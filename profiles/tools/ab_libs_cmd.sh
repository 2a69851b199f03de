# A/B of library variants built by build_variant.sh with an arbitrary timing command:
#   bash profiles/tools/ab_libs_cmd.sh "python profiles/tools/ab_grid2_spc.py 131072 0" PREV ...   (the in-tree library runs last)
CMD=$1; shift
cd generative-physics-informed-pde_b200
cp libgpde_b200.so build/libORIG.so
for v in "$@" ORIG; do cp build/lib$v.so libgpde_b200.so; echo "== $v"; (cd ..; timeout 200 $CMD 2>&1 | tail -3); done
cp build/libORIG.so libgpde_b200.so
