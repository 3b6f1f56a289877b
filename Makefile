# Builds everything in-tree (same steps as __graft_entry__.build()): the CUDA library, the host driver, the test-only
# oracle (+ oracle/_ref when /root/reference is present) and the test-only SIMT emulation of the kernels.
all:
	$(MAKE) -C psmc_b200/csrc
	$(MAKE) -C host
	$(MAKE) -C oracle
	$(MAKE) -C tests/emu
clean:
	$(MAKE) -C psmc_b200/csrc clean
	$(MAKE) -C host clean
	$(MAKE) -C tests/emu clean
.PHONY: all clean
