// NUMA placement of the host side of a device pipeline (SURVEY.md 8(e), 8(f)-1 "NUMA-aware pinned arenas").
//
// On a multi-socket host a GPU hangs off one socket's PCIe root: pinned memory on the other socket costs every
// frame copy a hop over the inter-socket link, and so do the feeder / output threads that touch it.  Linux places
// freshly allocated pages on the node of the CPU that first touches (here: pins) them, so binding the allocating
// thread to the GPU's node is enough -- no libnuma, no mempolicy calls.  On a single-node host (the B200 boxes of
// this pool: `nvidia-smi topo -m` reports NUMA affinity 0 for every GPU) every call here is a no-op.
// MP2V_NUMA=0 disables it.
#pragma once
#include <string>
#include <vector>

namespace mp2v {

// "0-3,8,10-11" -> {0,1,2,3,8,10,11}; malformed input -> empty
std::vector<int> parse_cpu_list(const std::string& s);

// NUMA node of a PCI device ("0000:3b:00.0", as cudaDeviceGetPCIBusId prints it) or -1 when the host has a single
// node, the kernel does not say, or MP2V_NUMA=0
int numa_node_of_pci_device(const std::string& bus_id);

// the CPUs of a node (empty: unknown)
std::vector<int> cpus_of_numa_node(int node);

// Binds the calling thread to the CPUs of `node` for the lifetime of the object and restores the previous affinity
// afterwards (node < 0: nothing happens).  Used around the pinned allocations of a device context.
class numa_scope_t {
public:
    explicit numa_scope_t(int node);
    ~numa_scope_t();
    numa_scope_t(const numa_scope_t&) = delete;
    numa_scope_t& operator=(const numa_scope_t&) = delete;
    bool bound() const { return bound_; }

private:
    bool bound_ = false;
    std::vector<unsigned long> saved_;      // the previous cpu_set_t, as words
};

// binds the calling thread to `node` for good (pipeline threads); false when nothing was done
bool bind_this_thread_to_numa_node(int node);

}  // namespace mp2v
