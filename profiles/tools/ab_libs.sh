# A/B timing of library variants built by build_variant.sh: bash profiles/tools/ab_libs.sh B V1 V1a ...
cd generative-physics-informed-pde_b200
cp libgpde_b200.so build/libORIG.so
for v in "$@"; do cp build/lib$v.so libgpde_b200.so; echo "== $v"; (cd ..; timeout 60 python profiles/tools/time_cfg3_vo.py 2>&1 | tail -1); done
cp build/libORIG.so libgpde_b200.so
