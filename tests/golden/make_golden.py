#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref/aiolos_ref_dump,
built by oracle/Makefile from /root/reference) on the reference's own decks.  Runs only in the dev container
(/root/reference does not exist on the GPU box); the resulting *.bin files are committed.

  <case>_init.bin     state right after the reference's constructor (input of aio_create_from_dump)
  <case>_k<K>.bin     state after exactly K steps of c_Sim::execute() from t=0 (trimmed to the compared fields)

usage: python tests/golden/make_golden.py [case ...]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from aiodump import read_dump, write_dump  # noqa: E402

RUN = "/tmp/aiolos_golden_run"
DUMP = os.path.join(ROOT, "oracle", "_ref", "aiolos_ref_dump")

# case -> (deck dir, par, spc or None, step counts)
CASES = {
    # BASELINE.json configs[0]: hydrostatic planet, 1 species, hydro only
    "c1_planet_spherical": ("runs_searth/", "planet_spherical.par", "mix3.spc", [1, 100, 2000]),
    # configs[1]: 6-species numeric friction (maintained twin of test2/cartesian_frictionMatrix6spc.par)
    "c2_friction_6spc": ("test_files/", "friction_6spc.par", "friction_6spc.spc", [1, 100, 1000]),
    # same deck made well-conditioned (T ~ 300 K, alpha*dt ~ 1): the shipped one starts at T ~ 1e-8 K where alpha*dt ~ 1e14
    # and the reference's own LU result depends on FMA contraction (README.md in this directory)
    "c2m_friction_6spc_mild": ("test_files/", "friction_6spc_mild.par", "friction_6spc.spc", [1, 100, 1000]),
    # gas + dust sound wave of the reference's own test suite: dust species (pressureless Riemann solver, advection.cpp:623-650),
    # constant collision model, problem 3 initial conditions, second order, periodic boundaries
    "d1_dustywave_nonstiff": ("test_files/", "dustywave_nonstiff.par", "hydrogen.spc", [1, 100, 800]),
    # configs[2]: 22-band full FLD
    "c3_searth_22b": ("test2/", "SEarth_opacities_1s22b.par", "irradiation.spc", [1, 20, 100, 1000]),
    # configs[3] made two-species at lower resolution (our variant): n-n friction + simple FLD with collisional T_i-T_j coupling
    "c5_dynamic3_2spc_collheat": ("runs_mdot_vs_euv_plot/", "dynamic3_2spc_collheat.par", None, [1, 100, 1000]),
    # the same with the full FLD solver: 3x3 blocks (J, T_1, T_2) with the collisional T-T coupling of radiation.cpp:589-615
    "c6_dynamic3_2spc_fullfld": ("runs_mdot_vs_euv_plot/", "dynamic3_2spc_fullfld.par", None, [1, 100, 1000]),
    # configs[3] with PARI_OPACITY_MODEL T and a reference .aiopa table (our variant): tabulated kappa(p, T) interpolation
    "c7_dynamic3_tabulated": ("runs_mdot_vs_euv_plot/", "dynamic3_tabulated.par", None, [1, 100, 1000]),
    # configs[3]: hydro + 2-band irradiation + simple FLD
    "c4_dynamic3_euv": ("runs_mdot_vs_euv_plot/", "dynamic3_euv1e3_C_1s2b_rerun.par", None, [1, 100, 1000, 3000]),
    # FRICTION_SOLVER 0, the closed-form two-species drag (source.cpp:249-330): constant and physical collision models
    "f1_friction_2spc_an": ("test_files/", "friction_2spc_an.par", "friction_2spc.spc", [1, 100, 500]),
    "f2_friction_2spc_phys_an": ("test_files/", "friction_2spc_phys_an.par", "friction_2spc.spc", [1, 100, 500]),
    # the three-species determinant formula of the same solver (source.cpp:332-400) on the well-conditioned friction deck
    # (our variant: friction_6spc_mild.par with 3 species and FRICTION_SOLVER 0)
    "f3_friction_3spc_an_mild": ("test_files/", "friction_3spc_an_mild.par", "friction_6spc.spc", [100, 1000]),
    # USE_PLANET_TEMPERATURE 1: the surface cell radiates at T_planet, slab model for T_planet (radiation.cpp:451-472,812-828)
    "t1_planetheating": ("runs_searth/", "SEarth_planetheating_1s1b.par", "mix3.spc", [100, 1000]),
    # ten species through the numeric friction solver (the rolled 16-species instantiation of the kernel); our variant of the mild
    # friction deck with the reference's 10-species file test2/mix_plt_radiationtest_multi.spc
    "f4_friction_10spc_mild": ("test_files/", "friction_10spc_mild.par", "mix_plt_radiationtest_multi.spc", [100, 1000]),
    # PARI_SELF_GRAV_SWITCH 1: the constructor's 50-iteration relaxation loop (init_and_bounds.cpp:645-662) and
    # update_mass_and_pot() every step / Heun stage (source.cpp:50-109); 2002 cells, full FLD
    "g1_saturn_selfgrav2": ("runs_schulik2020/", "saturn_1grey_1s1b_selfgrav2.par", "mix_grey1e-2.spc", [100, 400]),
    # DO_HYDRO 0: radiation-only relaxation (no hydro stages; at order 2 only the boundary conditions are applied,
    # advection.cpp:196-200; dt from the energy limiter, helpers.cpp:80-85), full FLD
    "n1_searth_nohydro": ("test2/", "SEarth_nohydro_1s1b.par", "mix_HHe.spc", [1, 100, 300]),
    # the reference's cartesian radiation test, 12 cells, run to its t_max (242 steps)
    "n2_radtest_1s1b": ("test2/", "cartesian_radiationtest_1spec1band.par", "mix_HHe.spc", [1, 100, 242]),
    # a long column (3054 cells > the 2046 one 256-thread CTA holds in shared memory): 1024 threads per column, work arrays in HBM
    "b1_hjwind_3054": ("runs_searth/", "hjwind_dgrey_1s1b.par", "mix3.spc", [100, 600]),
    # configs[3] at three times its resolution (2987 cells; our variant): hydro + irradiation + simple FLD on a long column
    "b2_dynamic3_3k": ("runs_mdot_vs_euv_plot/", "dynamic3_euv_3k.par", None, [100, 600]),
    # PBC_model/dusty_wind.par at 1000 instead of 4000 cells per decade (our variant), with the reference linked against its own
    # problems/dusty_wind.cpp (oracle/_ref/aiolos_ref_dump_dusty_wind): user initial conditions + user boundary hooks on both
    # sides, gas + dust, numeric friction, opacity model D, full FLD with 3x3 blocks
    "p1_dusty_wind": ("PBC_model/", "dusty_wind_1k.par", "silicate.spc", [1, 20, 100, 1000], "dusty_wind"),
    # the shipped deck itself, unmodified (4002 cells): plugin + long column; only the 300-step state is kept (no init dump:
    # plugin decks are created from files)
    "p2_dusty_wind_full": ("PBC_model/", "dusty_wind.par", "silicate.spc", [300], "dusty_wind", "noinit"),
    # configs[0] with the wave driver switched on (WAVE_USE 1: add_wave() sets the momentum of the two left ghost cells every stage,
    # init_and_bounds.cpp:1982-1987, advection.cpp:425-426; our variant -- the shipped wave decks test2/*wave*.par are unit-less
    # legacy decks whose ghost cells hold cancellation noise, profiles/r1l_parity_sweep.md)
    "w1_planet_wave": ("runs_searth/", "planet_spherical_wave.par", "mix3.spc", [1, 100, 1000]),
    # configs[0] with the giantplanet boundary on the left (PARI_BOUND_TYPE_LEFT 5; our variant: no shipped deck selects it)
    "gp1_planet_giantplanet": ("runs_searth/", "planet_spherical_giantplanet.par", "mix3.spc", [1, 100, 1000]),
    # c5 (two species, n-n friction, simple FLD + collisional heat exchange) with KZZ_INIT 1e-10 and a second species of 16 amu
    # (our variant; no shipped deck sets KZZ_INIT): update_kzz_and_gravpot's homopause scaling of phi_s (source.cpp:174-219), every step
    "z1_dynamic3_2spc_kzz": ("runs_mdot_vs_euv_plot/", "dynamic3_2spc_kzz.par", None, [1, 100, 1000]),
    # the shipped C2Ray deck c2ray/hd209458b_xrays_dynamicGreyHeating_3s2b.par (atomic H, protons, electrons; PARI_USE_RADIATION 0)
    # with the ion / electron start fractions of its species file raised from 1e-30 to 1e-6 (our variant mix_xrays_ions1e-6.spc:
    # with 1e-30 the unmodified reference stops in its first step, "Root must be bracketed", because the proton density
    # underflows to 0 and the heat-exchange matrix becomes infinite).  h0: the same with PHOTOCHEM_LEVEL 0 -- three charged /
    # neutral species through the numeric friction solver with the physical collision model (i-n, e-n, i-e classes).
    "h0_hd209458b_nophotochem": ("c2ray/", "hd209458b_nophotochem.par", "mix_xrays_ions1e-6.spc", [1, 100, 1000]),
    # h1: PHOTOCHEM_LEVEL 1, the C2Ray ionisation + heating sweep (photochem.cpp:430-818)
    "h1_hd209458b_c2ray": ("c2ray/", "hd209458b_xrays_dynamicGreyHeating_3s2b.par", "mix_xrays_ions1e-6.spc", [1, 100, 400, 800]),
}

KEEP_FINAL = ("dt_sequence", "t_max", "t_max_deck", "steps", "globalTime", "dt", "cflfactor", "timestep_rad2", "cfl_step",
              "phi", "Jrad_FLD", "dS_band", "total_opacity", "T_planet")
KEEP_SPECIES = ("u1", "u2", "u3", "temperature", "speed", "pres", "eint_last", "dS", "sound_speed", "phi_s")


def run_case(name):
    d, par, spc, steps = CASES[name][:4]
    dump = DUMP + ("_" + CASES[name][4] if len(CASES[name]) > 4 else "")      # plugin variant of the oracle
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "make_rundir.sh"), RUN], check=True, stdout=subprocess.DEVNULL)
    for f in os.listdir(os.path.join(HERE, "decks")):        # our own deck variants live next to the reference's test decks
        subprocess.run(["cp", os.path.join(HERE, "decks", f), os.path.join(RUN, "test_files", f)], check=True)
        subprocess.run(["cp", os.path.join(HERE, "decks", f), os.path.join(RUN, "runs_mdot_vs_euv_plot", f)], check=True)
        subprocess.run(["cp", os.path.join(HERE, "decks", f), os.path.join(RUN, "PBC_model", f)], check=True)
        subprocess.run(["cp", os.path.join(HERE, "decks", f), os.path.join(RUN, "c2ray", f)], check=True)
        subprocess.run(["cp", os.path.join(HERE, "decks", f), os.path.join(RUN, "runs_searth", f)], check=True)
    for k in steps:
        out = os.path.join(RUN, f"{name}_k{k}")
        cmd = [dump, "-dir", d, "-par", par, "-nsteps", str(k), "-out", out]
        if spc:
            cmd += ["-spc", spc]
        with open(out + ".log", "w") as log:
            subprocess.run(cmd, cwd=RUN, stdin=subprocess.DEVNULL, stdout=log, stderr=subprocess.STDOUT, check=True)
        fin = read_dump(out + "_final.bin")
        assert int(fin["steps"][0]) == k, (name, k, fin["steps"])
        keep = {key: v for key, v in fin.items()
                if key in KEEP_FINAL or (key[0] == "s" and "." in key and key.split(".", 1)[1] in KEEP_SPECIES)}
        write_dump(os.path.join(HERE, f"{name}_k{k}.bin"), keep)
        if k == steps[0] and "noinit" not in CASES[name]:
            init = read_dump(out + "_init.bin")
            write_dump(os.path.join(HERE, f"{name}_init.bin"), init)
        print(f"{name}: k={k} t={fin['globalTime'][0]:.6e} dt={fin['dt'][0]:.6e}")


def run_nofma():
    """configs[1] as shipped with the reference built WITHOUT FMA contraction (`make -C oracle nofma`): the reference's own
    scatter on that deck (README.md in this directory)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "nofma", "aiolos_ref_dump")
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "nofma"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["bash", os.path.join(ROOT, "oracle", "make_rundir.sh"), RUN], check=True, stdout=subprocess.DEVNULL)
    for k in (100, 1000):
        out = os.path.join(RUN, f"c2nofma_k{k}")
        subprocess.run([exe, "-dir", "test_files/", "-par", "friction_6spc.par", "-spc", "friction_6spc.spc", "-nsteps", str(k), "-out", out],
                       cwd=RUN, stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT, check=True)
        fin = read_dump(out + "_final.bin")
        keep = {key: v for key, v in fin.items() if key in ("dt_sequence", "steps", "globalTime", "dt")
                or (key[0] == "s" and "." in key and key.split(".", 1)[1] in ("u1", "u2", "u3", "sound_speed"))}
        write_dump(os.path.join(HERE, f"c2_friction_6spc_nofma_k{k}.bin"), keep)
        print(f"c2_friction_6spc (no FMA contraction): k={k} steps={int(fin['steps'][0])} t={fin['globalTime'][0]:.6e}")


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        if c == "nofma":
            run_nofma()
        else:
            run_case(c)
