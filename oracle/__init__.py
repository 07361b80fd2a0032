"""CPU oracle for the rasteriser hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product
package (gaussian_splatting_3d_b200), which has no CPU fallback."""
