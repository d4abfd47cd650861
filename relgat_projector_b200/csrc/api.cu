// ABI bookkeeping for librelgat_b200.so (see include/relgat_b200.h).
extern "C" int relgat_abi_version(void) { return 13; }
