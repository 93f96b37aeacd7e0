"""mdacp_b200 -- B200-native (sm_100a) replacement for the data-parallel hot path of kaityo256/mdacp.

capi    : ctypes binding of the C-ABI library (include/mdacp_b200.h), no CPU fallback
driver  : host-side mirror of the reference driver surface (MDManager / ParaInfo / SimulationInfo)
"""
__version__ = "0.1.0"
