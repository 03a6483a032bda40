"""CPU oracle for the RD-map + OS-CFAR path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy, a little pure Python for the literal
stream model) of the reference FPGA chain
    window -> range FFT -> corner turn -> (MTI) -> window -> Doppler FFT
           -> magnitude -> 2D OS-CFAR -> detection list
as written in /root/reference/rtl/src/*.vhd.  It exists only so that
    * tests/                       (the checker of the CUDA path),
    * __graft_entry__.smoke()      (one small GPU-vs-oracle check),
    * bench.py's cpu_baseline leg  (and `bench.py --impl reference`)
have something to compare against / time.  Nothing under the product package
`fpga-fmcw-radar-processor_b200/` imports it, and the product never falls back
to it: if the CUDA library is missing the product raises.

Pinning status (see DESIGN.md "Oracle"):
  * rtl_fixed.magnitude     pinned by the 17 known-answer vectors of
                            rtl/src/tb_magnitude_calc.vhd:49-73 (exact).
  * rtl_fixed.corner_turn   pinned by rtl/src/tb_corner_turner.vhd:146-186.
  * rtl_fixed.window        pinned by the range/symmetry/saturation checks of
                            rtl/src/tb_window_multiplier.vhd:83-108,185-240.
  * rtl_fixed.mti           pinned by rtl/src/tb_doppler_notch.vhd:117-180.
  * rtl_fixed.os_cfar_2d*   pinned by the detection-count checks of
                            rtl/src/tb_os_cfar_2d.vhd:123-209 and by a literal
                            clock-by-clock model of os_cfar_2d.vhd
                            (rtl_fixed.os_cfar_2d_stream).
  * float_path (FFT stages) PARITY UNPINNED against the FPGA arithmetic: the
                            FFTs are closed Xilinx xfft v9.1 IP whose RTL is
                            absent and no reference test asserts their output.
                            Pinned only structurally: forward DFT, natural
                            order, golden_input_chirp.txt peak at bin 290,
                            radar_output.txt target rows {99..101},{499..501}.
  * rtl_fixed.bfp_fft, process_cpi_fixed
                            PARITY UNPINNED for the two FFT stages: a
                            SELF-DEFINED 16-bit block-floating-point transform
                            (the IP's word widths and convergent rounding, not
                            its netlist) chained with the pinned stages above;
                            librdp's arith = FIXED16 is bit-exact against it
                            ("self-consistent, not Xilinx-exact").
  * float_path.peak_group, combine = "nci" / "beams"
                            not in the reference at all: the
                            rules librdp's optional stages implement, stated here
                            so that they can be tested.
"""
from . import rtl_fixed, float_path, stimulus  # noqa: F401
