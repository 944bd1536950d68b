"""Named parity cases shared by the golden generator and the tests (sizes the CPU oracle finishes in seconds)."""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ISO_MESH = os.path.join(ROOT, "config", "mesh", "unit-square-iso.mesh")
BC2_FILE = {1: -0.5, 2: 0.5}                       # the shipped two-triangle mesh: attr 2 = top
BC2 = {1: -0.5, 2: -0.5, 3: 0.5, 4: -0.5}          # builtin 2-D meshes: bottom,right,top,left
BC3 = {1: -0.5, 2: -0.5, 3: -0.5, 4: -0.5, 5: -0.5, 6: 0.5}  # builtin 3-D meshes: top (z=1) hot

# name -> dict(mesh, order, refine, angle_dim, polar, azimuth, polar_gauss, azimuth_gauss, bcs, num_spectral, iters)
CASES = {
    # C0: the reference's shipped default run (config/config.yaml + si.yaml), 101 iterations
    "c0_default": dict(mesh=ISO_MESH, order=1, refine=0, angle_dim=2, polar=24, azimuth=24, polar_gauss=True, azimuth_gauss=True,
                       bcs=BC2_FILE, num_spectral=20, iters=101),
    # the run behind the shipped sweep_dim3 log: refined once, order 2, 24x24 Gauss directions on a 2-D mesh
    "c0_refined_dim3": dict(mesh=ISO_MESH, order=2, refine=1, angle_dim=3, polar=24, azimuth=24, polar_gauss=True, azimuth_gauss=True,
                            bcs=BC2_FILE, num_spectral=1, iters=2),
    # C1: builtin 8x8x2 triangles, order 1, 16 uniform azimuth points, "gray" = 2 bands
    "c1_tri_p1": dict(mesh="unit-square-tri", order=1, refine=0, angle_dim=2, polar=1, azimuth=16, polar_gauss=True, azimuth_gauss=False,
                      bcs=BC2, num_spectral=1, iters=6),
    "tri_p3_si": dict(mesh="unit-square-tri", order=3, refine=0, angle_dim=2, polar=1, azimuth=8, polar_gauss=True, azimuth_gauss=True,
                      bcs=BC2, num_spectral=4, iters=4),
    "quad_p2": dict(mesh="unit-square-quad", order=2, refine=0, angle_dim=2, polar=1, azimuth=8, polar_gauss=True, azimuth_gauss=False,
                    bcs=BC2, num_spectral=2, iters=4),
    "quad_p4": dict(mesh="unit-square-quad", order=4, refine=0, angle_dim=2, polar=1, azimuth=8, polar_gauss=True, azimuth_gauss=False,
                    bcs=BC2, num_spectral=1, iters=3),
    "tet_p1": dict(mesh="unit-cube-tet", order=1, refine=0, angle_dim=3, polar=2, azimuth=4, polar_gauss=True, azimuth_gauss=False,
                   bcs=BC3, num_spectral=1, iters=4),
    "tet_p2": dict(mesh="unit-cube-tet", order=2, refine=0, angle_dim=3, polar=4, azimuth=4, polar_gauss=True, azimuth_gauss=False,
                   bcs=BC3, num_spectral=2, iters=3),
    "hex_p1": dict(mesh="unit-cube-hex", order=1, refine=0, angle_dim=3, polar=2, azimuth=8, polar_gauss=True, azimuth_gauss=False,
                   bcs=BC3, num_spectral=1, iters=4),
    # small sibling of BASELINE config C2 (hex, order 2, 8 Gauss polar x 16 uniform azimuth, 2 bands) on the unrefined 4^3 mesh
    "c2_hex_p2_small": dict(mesh="unit-cube-hex", order=2, refine=0, angle_dim=3, polar=8, azimuth=16, polar_gauss=True, azimuth_gauss=False,
                            bcs=BC3, num_spectral=1, iters=3),
    # ndof = 64 (hex order 3): exercises the generic kernel
    "hex_p3": dict(mesh="unit-cube-hex", order=3, refine=0, angle_dim=3, polar=2, azimuth=4, polar_gauss=True, azimuth_gauss=False,
                   bcs=BC3, num_spectral=1, iters=2),
    # 2-D mesh with 3-D angles on the 2-D builtin
    "quad_p1_dim3": dict(mesh="unit-square-quad", order=1, refine=0, angle_dim=3, polar=4, azimuth=8, polar_gauss=True, azimuth_gauss=False,
                         bcs=BC2, num_spectral=1, iters=4),
}


def host_case(name, **override):
    """Build the problem image of a named case through the host layer."""
    import pbte_b200 as pb
    c = dict(CASES[name])
    c.update(override)
    return pb.HostCase(c["mesh"], c["order"], c["refine"], c["angle_dim"], c["polar"], c["azimuth"], c["polar_gauss"], c["azimuth_gauss"],
                       c["bcs"], face_rule=c.get("face_rule", 0), num_spectral=c["num_spectral"])
