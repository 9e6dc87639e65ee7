"""Minimal astropy stand-in (TEST INFRASTRUCTURE ONLY).

astropy is not installable in this image.  The reference uses it only for ten CGS
constants / unit conversions (LeMoC/LeHaMoC_f.py:17-31) and for BlackBody when
BB_flag=1 (LeMoC/LeMoC.py:144-146).  Values are the astropy 5.0.4 ones pinned by the
reference's requirements.yml (CODATA 2018 / IAU 2015).  Used only by
oracle/make_golden.py to run the unmodified reference from /root/reference in the build
container; nothing in the product imports it.
"""
