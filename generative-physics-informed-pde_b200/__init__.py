"""B200-native physics layer of pkmtum/generative-physics-informed-pde (ROM + virtual observables).

Import as ``gpde_b200`` (the repo-root shim ``gpde_b200.py`` maps that name onto this directory,
whose on-disk name is not a valid Python identifier).

    from gpde_b200.ROM import ROM                          # bottleneck/ROM.py
    from gpde_b200.components import ReducedOrderModelOperator   # bottleneck/components.py:260-323
    from gpde_b200 import VirtualObservables               # bottleneck/VirtualObservables.py
    from gpde_b200.physics import setup_physics            # FEniCS-free setup exporter

Hot-path modules need libgpde_b200.so (build.py); there is no CPU fallback.
"""
__version__ = "0.1.0"
