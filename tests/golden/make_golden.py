"""Generates tests/golden/*.npz by running the UNMODIFIED reference classes (bottleneck/ROM.py,
bottleneck/components.py, bottleneck/VirtualObservables.py from /root/reference) under the stub-dolfin
shim (oracle/ref_shim.py), fed with the constants of the restated P1 assembler (oracle/fem_p1.py).

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The fixtures pin (a) the oracle restatements in oracle/rom_ref.py / oracle/vo_ref.py and (b) through
them the CUDA path.  They do NOT pin the FEniCS boundary (mesh/dof numbering, assembled constants):
FEniCS is not installable here -- "parity unpinned" for that part, see DESIGN.md.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fem_p1, ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synthetic_inputs(P, B, kind, seed, ell):
    rng = np.random.RandomState(seed)
    nx_f, ny_f = P['nx_fom'], P['ny_fom']
    img = fem_p1.sample_log_field(ny_f, nx_f, 0.4, 0.8, ell, B, rng)          # fine log-conductivity images
    X_DG = fem_p1.image_to_function(img, P['pixel_of_cell_fom'])               # [B, E_f]
    # coarse log-conductivity: mean of the fine log field over each coarse cell
    E = len(P['cells_rom'])
    mid = P['coords_fom'][P['cells_fom']].mean(axis=1)
    nx, ny = P['nx_rom'], P['ny_rom']
    sx = np.minimum((mid[:, 0] * nx).astype(int), nx - 1)
    sy = np.minimum((mid[:, 1] * ny).astype(int), ny - 1)
    upper = (mid[:, 1] * ny - sy) > (mid[:, 0] * nx - sx)
    owner = 2 * (sy * nx + sx) + upper.astype(int)
    logX = np.stack([X_DG[:, owner == e].mean(axis=1) for e in range(E)], axis=1)
    coef = rng.uniform(-0.5, 0.5, size=(B, 4))
    g_rom = np.stack([fem_p1.dirichlet_left_right(P['coords_rom'], kind, coef[b])[1] for b in range(B)])
    g_fom = np.stack([fem_p1.dirichlet_left_right(P['coords_fom'], kind, coef[b])[1] for b in range(B)])
    F = fem_p1.full_F_with_applied_bc(len(P['coords_rom']), P['bc_dofs_rom'], g_rom)
    gbar = rng.normal(size=(B, len(P['coords_rom'])))
    return dict(img=img, X_DG=X_DG, logX=logX, bc_coef=coef, g_rom=g_rom, g_fom=g_fom, F=F, gbar_u=gbar)


def rom_case(ref, name, nx, refines, B, kind, seed, ell):
    P = fem_p1.build_problem(nx, nx, refines)
    inp = synthetic_inputs(P, B, kind, seed, ell)
    ROM = ref['ROM'].ROM
    Operator = ref['components'].ReducedOrderModelOperator
    phys = ref_shim.PhysicsLike(P['bc_dofs_rom'], P['free_dofs_rom'], len(P['cells_rom']))
    M = torch.tensor(P['M'], dtype=torch.double)
    rom = ROM(phys, M, torch.double, torch.device('cpu'))
    W = torch.tensor(P['W'], dtype=torch.double)
    op = Operator(rom, W, dtype=torch.double, device=torch.device('cpu'))

    logX = torch.tensor(inp['logX'], requires_grad=True)
    F = torch.tensor(inp['F'], requires_grad=True)
    gbar = torch.tensor(inp['gbar_u'])
    # ROM.__call__ on conductivities, exactly as components.py:298 feeds it
    x = torch.exp(logX) + 1e-8
    u, K = rom(x, F, ReturnStiffness=True)
    u.backward(gbar)
    out = dict(u=u.detach().numpy(), K=K.detach().numpy(), grad_logX=logX.grad.numpy().copy(),
               grad_F=F.grad.numpy().copy())
    # gradient w.r.t. the conductivities themselves (ROM used directly)
    xd = x.detach().clone().requires_grad_(True)
    rom(xd, F.detach()).backward(gbar)
    out['grad_x'] = xd.grad.numpy().copy()
    # operator: mean and the gradient of sum(w * mu_y) w.r.t. effprop
    eff = torch.tensor(inp['logX'], requires_grad=True)
    mu_y, ls_y = op.forward(eff, F.detach())
    wy = torch.tensor(np.random.RandomState(seed + 1).normal(size=tuple(mu_y.shape)))
    (mu_y * wy).sum().backward()
    out.update(mu_y=mu_y.detach().numpy(), gbar_y=wy.numpy(), grad_effprop=eff.grad.numpy().copy(),
               logsigmas_shape=np.array(ls_y.shape))
    consts = dict(M=P['M'], W=P['W'], bc_dofs_rom=P['bc_dofs_rom'], free_dofs_rom=P['free_dofs_rom'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), nx=nx, refines=refines, kind=kind,
                        **{'in_' + k: v for k, v in inp.items() if k in ('logX', 'F', 'gbar_u', 'bc_coef')},
                        **{'out_' + k: v for k, v in out.items()}, **{'const_' + k: v for k, v in consts.items()})
    print(name, 'u', out['u'].shape, 'max|u|', np.abs(out['u']).max())


def vo_case(ref, name, nx, refines, N, kind, seed, ell, n_rbf):
    VOm = ref['VirtualObservables']
    P = fem_p1.build_problem(nx, nx, refines)
    inp = synthetic_inputs(P, N, kind, seed, ell)
    rng = np.random.RandomState(seed + 7)
    d = len(P['free_dofs_fom'])

    class BC(object):   # what assemble_system / samplers read from a boundary condition
        def __init__(self, g):
            self.g = g

    def assemble(x, bc):
        return fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], x, P['bc_dofs_fom'], bc.g,
                                           P['free_dofs_fom'])

    phys = ref_shim.PhysicsLike(P['bc_dofs_fom'], P['free_dofs_fom'], len(P['cells_fom']), assemble)
    centres = rng.uniform(size=(n_rbf, 2))
    V = np.hstack([P['W'], fem_p1.rbf_columns(P['coords_fom'], P['free_dofs_fom'], centres, 0.1)])
    m = V.shape[1]
    mask = -np.ones(m)
    mask[-2:] = 1.0    # two learnable-precision observables so that update_vo_precision does work

    class FixedSampler(VOm.BaseSampler):   # constant sampler around the reference's own weak-Galerkin code
        def __init__(self, qp):
            super().__init__(qp)
            self._GA = qp.construct_querry_weak_galerkin(V)
        m = property(lambda self: V.shape[1])
        is_constant = property(lambda self: True)
        precision_mask = property(lambda self: mask)

        def sample(self):
            return self._GA

    dev = torch.device('cpu')
    qps = [VOm.QuerryPoint(phys, inp['X_DG'][n], BC(inp['g_fom'][n])) for n in range(N)]
    qpe = VOm.QuerryPointEnsemble(qps)
    qe = VOm.QuerryEnsemble([VOm.LinearQuerry(qp, FixedSampler(qp), torch.double, dev) for qp in qps], torch.double, dev)
    ens = VOm.VirtualObservablesEnsemble(qpe, qe, torch.double, dev)

    Y = rng.normal(size=(N, d)) * 0.1 + (P['W'] @ P['coords_rom'][:, 0])[None]       # something like a solution
    G1 = torch.tensor(Y)
    PREC1 = torch.tensor(rng.uniform(50.0, 200.0, size=(N, d)))
    ens.update(G1, PREC1, 0)
    mean1, vars1 = ens.mean.numpy().copy(), ens.vars.numpy().copy()
    G2 = torch.tensor(mean1 + 0.01 * rng.normal(size=(N, d)))
    PREC2 = torch.tensor(rng.uniform(50.0, 200.0, size=(N, d)))
    ens.update(G2, PREC2, 1)      # this one runs update_vo_precision (VirtualObservables.py:971-998)
    out = dict(
        Gamma=np.stack([q.Gamma.numpy() for q in qe]), alpha=np.stack([q.alpha.numpy() for q in qe]),
        residual=np.stack([q.Gamma.numpy() @ Y[n] - q.alpha.numpy() for n, q in enumerate(qe)]),
        mean1=mean1, vars1=vars1, mean2=ens.mean.numpy().copy(), vars2=ens.vars.numpy().copy(),
        prec_beta=ens._prec_beta.numpy().copy(), mean_vo_variances=ens._mean_vo_variances.numpy().copy(),
    )
    np.savez_compressed(os.path.join(OUT, name + '.npz'), nx=nx, refines=refines, kind=kind,
                        in_X_DG=inp['X_DG'], in_g_fom=inp['g_fom'], in_bc_coef=inp['bc_coef'], in_V=V, in_mask=mask,
                        in_Y=Y, in_PREC1=PREC1.numpy(), in_G2=G2.numpy(), in_PREC2=PREC2.numpy(),
                        **{'out_' + k: v for k, v in out.items()})
    print(name, 'Gamma', out['Gamma'].shape, 'max|r|', np.abs(out['residual']).max())


def energy_case(ref, name, nx, refines, N, kind, seed, ell, m_sub, n_it):
    """EnergyVirtualObservable(sEnsemble) of the reference (VirtualObservables.py:672-788, 1001-1037) run under the
    shim with a deterministic sampler (a fixed sequence of RBF weighting matrices) and a linear temperature schedule."""
    VOm = ref['VirtualObservables']
    P = fem_p1.build_problem(nx, nx, refines)
    inp = synthetic_inputs(P, N, kind, seed, ell)
    rng = np.random.RandomState(seed + 200)
    d = len(P['free_dofs_fom'])

    class BC(object):
        def __init__(self, g):
            self.g = g

    def assemble(x, bc):
        return fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], x, P['bc_dofs_fom'], bc.g,
                                           P['free_dofs_fom'])

    phys = ref_shim.PhysicsLike(P['bc_dofs_fom'], P['free_dofs_fom'], len(P['cells_fom']), assemble)
    n_updates = 2
    V_seq = np.stack([fem_p1.rbf_columns(P['coords_fom'], P['free_dofs_fom'], rng.uniform(size=(m_sub, 2)), 0.25)
                      for _ in range(n_updates * N * n_it)])            # consumed in call order

    class SequenceSampler(VOm.BaseSampler):
        calls = 0

        def _sample(self):
            V = V_seq[SequenceSampler.calls]
            SequenceSampler.calls += 1
            return V

    dev = torch.device('cpu')
    qps = [VOm.QuerryPoint(phys, inp['X_DG'][n], BC(inp['g_fom'][n])) for n in range(N)]
    qpe = VOm.QuerryPointEnsemble(qps)
    ens = VOm.EnergyVirtualObservablesEnsemble(qpe, n_it, SequenceSampler(qps[0]), torch.double, dev)
    ens.set_linear_temperature_schedule(T_init=1.0, T_final=1e-2, num_steps=4)
    G = rng.normal(size=(n_updates, N, d)) * 0.1 + (P['W'] @ P['coords_rom'][:, 0])[None, None]
    PREC = rng.uniform(50.0, 200.0, size=(n_updates, N, d))
    means, varss, temps = [], [], []
    for it in range(n_updates):
        ens.update(torch.tensor(G[it]), torch.tensor(PREC[it]), it)
        means.append(ens.mean.numpy().copy()); varss.append(ens.vars.numpy().copy()); temps.append(ens[0].temperature)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), nx=nx, refines=refines, kind=kind, n_it=n_it,
                        in_X_DG=inp['X_DG'], in_g_fom=inp['g_fom'], in_bc_coef=inp['bc_coef'], in_V_seq=V_seq,
                        in_G=G, in_PREC=PREC, out_mean=np.stack(means), out_vars=np.stack(varss),
                        out_temperature=np.array(temps))
    print(name, 'energy VO', np.stack(means).shape, 'T', temps)


class _DataSet(object):
    """What GenerativeModel reads from a data set (utils/data.py is outside the hot path): .get(key) and .N."""

    def __init__(self, **tensors):
        self._t = tensors
        self.N = next(iter(tensors.values())).shape[0]

    def get(self, key, random_subset=None):
        return self._t[key]

    def __bool__(self):
        return True


def elbo_case(ref, name, nx, refines, Ns, Nvo, kind, seed, ell, dim_latent):
    """The UNMODIFIED GenerativeModel.elbo of the reference (bottleneck/generative.py:247-287, 352-392, 456-500) on a
    supervised and a virtual-observable data set, with the reference's own ROM / ReducedOrderModelOperator /
    VirtualObservablesEnsemble / VariationalApproximation / EffectivePropertyMap; the CNN decoder is replaced by
    oracle/elbo_ref.TinyDecoder.  Stores ELBO, every parameter gradient, the initial parameters and the noise draws."""
    import importlib
    from oracle import elbo_ref
    gen = importlib.import_module("bottleneck.generative")
    comp, VOm = ref['components'], ref['VirtualObservables']
    P = fem_p1.build_problem(nx, nx, refines)
    dt, dev = torch.double, torch.device('cpu')
    sup = synthetic_inputs(P, Ns, kind, seed, ell)
    vo = synthetic_inputs(P, Nvo, kind, seed + 1, ell)
    rng = np.random.RandomState(seed + 50)
    d, n_pix = len(P['free_dofs_fom']), P['nx_fom'] * P['ny_fom']

    phys_rom = ref_shim.PhysicsLike(P['bc_dofs_rom'], P['free_dofs_rom'], len(P['cells_rom']))
    rom = ref['ROM'].ROM(phys_rom, torch.tensor(P['M'], dtype=dt), dt, dev)
    g = comp.ReducedOrderModelOperator(rom, torch.tensor(P['W'], dtype=dt), dtype=dt, device=dev)
    torch.manual_seed(seed)
    f = elbo_ref.TinyDecoder(dim_latent, n_pix, dt, dev)
    gp = comp.EffectivePropertyMap(dim_latent, len(P['cells_rom']), num_hidden_layers=0, independent_X=True, dtype=dt, device=dev)
    model = gen.GenerativeModel(f=f, g=g, gp=gp, dtype=dt, device=dev)

    # supervised labels: the fine solution of each field (oracle assembler), restricted to the free dofs
    import scipy.sparse.linalg as spla
    Y = np.zeros((Ns, d))
    for b in range(Ns):
        K, fe = fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], np.exp(sup['X_DG'][b]), P['bc_dofs_fom'],
                                            sup['g_fom'][b], P['free_dofs_fom'])
        Y[b] = spla.spsolve(K.tocsc(), fe)
    ds_s = _DataSet(X=torch.tensor(sup['img'].reshape(Ns, -1)), Y=torch.tensor(Y), F_ROM_BC=torch.tensor(sup['F']))
    ds_v = _DataSet(X=torch.tensor(vo['img'].reshape(Nvo, -1)), F_ROM_BC=torch.tensor(vo['F']))

    # virtual observables: the reference ensemble with V = W (coarse-grained residuals, infinite precision)
    class BC(object):
        def __init__(self, gv):
            self.g = gv

    def assemble(x, bc):
        return fem_p1.assemble_system_free(P['coords_fom'], P['cells_fom'], x, P['bc_dofs_fom'], bc.g, P['free_dofs_fom'])

    phys_fom = ref_shim.PhysicsLike(P['bc_dofs_fom'], P['free_dofs_fom'], len(P['cells_fom']), assemble)
    V = P['W']

    class FixedSampler(VOm.BaseSampler):
        def __init__(self, qp):
            super().__init__(qp)
            self._GA = qp.construct_querry_weak_galerkin(V)
        m = property(lambda self: V.shape[1])
        is_constant = property(lambda self: True)
        precision_mask = property(lambda self: -np.ones(V.shape[1]))

        def sample(self):
            return self._GA

    qps = [VOm.QuerryPoint(phys_fom, vo['X_DG'][n], BC(vo['g_fom'][n])) for n in range(Nvo)]
    qpe = VOm.QuerryPointEnsemble(qps)
    qe = VOm.QuerryEnsemble([VOm.LinearQuerry(qp, FixedSampler(qp), dt, dev) for qp in qps], dt, dev)
    ens = VOm.VirtualObservablesEnsemble(qpe, qe, dt, dev)
    G = rng.normal(size=(Nvo, d)) * 0.1 + (P['W'] @ P['coords_rom'][:, 0])[None]
    PREC = rng.uniform(20.0, 100.0, size=(Nvo, d))
    ens.update(torch.tensor(G), torch.tensor(PREC), 0)

    model.register_datasets(dict(supervised=ds_s, vo=ds_v), VO=ens)
    with torch.no_grad():      # non-trivial variational parameters (the reference initialises them to zero)
        for key, N in (("supervised", Ns), ("vo", Nvo)):
            model.q_z[key]._mean.copy_(torch.tensor(rng.normal(size=(N, dim_latent)) * 0.3))
            model.q_z[key]._logsigma.copy_(torch.tensor(rng.normal(size=(N, dim_latent)) * 0.1 - 1.0))
            src = sup if key == "supervised" else vo
            model.q_X[key]._mean.copy_(torch.tensor(src['logX'] + 0.05 * rng.normal(size=src['logX'].shape)))
            model.q_X[key]._logsigma.copy_(torch.tensor(rng.normal(size=src['logX'].shape) * 0.1 - 2.0))
        g.logsigmas_y.copy_(torch.tensor(rng.normal(size=d) * 0.1 - 1.5))
    params = elbo_ref.named_parameters(f, gp, g, model.q_z, model.q_X)
    init = {k: p.detach().numpy().copy() for k, p in params.items()}
    with elbo_ref.NoiseTape() as tape:
        value = model.elbo(step=0)
    value.backward()
    grads = {k: p.grad.numpy().copy() for k, p in params.items()}
    np.savez_compressed(
        os.path.join(OUT, name + '.npz'), nx=nx, refines=refines, kind=kind, dim_latent=dim_latent,
        in_sup_img=sup['img'], in_sup_Y=Y, in_sup_F=sup['F'], in_sup_bc_coef=sup['bc_coef'],
        in_vo_img=vo['img'], in_vo_X_DG=vo['X_DG'], in_vo_F=vo['F'], in_vo_bc_coef=vo['bc_coef'], in_vo_g_fom=vo['g_fom'],
        in_vo_G=G, in_vo_PREC=PREC, out_vo_mean=ens.mean.numpy(), out_vo_logsigma=ens.logsigma.numpy(),
        out_elbo=np.array(value.item()), n_noise=len(tape.draws),
        **{'noise_%d' % i: dr.numpy() for i, dr in enumerate(tape.draws)},
        **{'init_' + k: v for k, v in init.items()}, **{'grad_' + k: v for k, v in grads.items()})
    print(name, 'ELBO', value.item(), 'params', len(params), 'noise draws', len(tape.draws))


def main():
    if not ref_shim.available():
        raise SystemExit('reference tree not found; fixtures can only be regenerated in the build container')
    ref = ref_shim.load()
    torch.manual_seed(0)
    np.random.seed(0)
    only = sys.argv[1] if len(sys.argv) > 1 else None            # regenerate one family: rom | vo | energy | elbo
    if only in (None, 'rom'):
        rom_case(ref, 'rom_4x4_ndp', 4, 3, 16, 'NDP', 0, 0.15)      # example.ipynb / highres32 shapes
        rom_case(ref, 'rom_8x8_nd', 8, 2, 8, 'ND', 1, 0.08)         # highres coarse mesh (fine mesh irrelevant here)
    if only in (None, 'vo'):
        vo_case(ref, 'vo_4x4_32_ndp', 4, 3, 3, 'NDP', 2, 0.15, n_rbf=5)
        vo_case(ref, 'vo_2x2_8_nd', 2, 2, 4, 'ND', 3, 0.3, n_rbf=3)
    if only in (None, 'energy'):
        energy_case(ref, 'energy_2x2_16_ndp', 2, 3, 3, 'NDP', 5, 0.3, m_sub=6, n_it=3)
    if only in (None, 'elbo'):
        elbo_case(ref, 'elbo_4x4_16_ndp', 4, 2, 6, 5, 'NDP', 11, 0.2, dim_latent=4)


if __name__ == '__main__':
    main()
