#include "numa.h"

#include <dirent.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace mp2v {

std::vector<int> parse_cpu_list(const std::string& s) {
    std::vector<int> out;
    size_t i = 0;
    const size_t n = s.size();
    auto number = [&](int& v) {
        if (i >= n || !isdigit((unsigned char)s[i])) return false;
        long x = 0;
        while (i < n && isdigit((unsigned char)s[i])) { x = x * 10 + (s[i] - '0'); if (x > (1 << 20)) return false; i++; }
        v = (int)x;
        return true;
    };
    while (i < n) {
        while (i < n && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t')) i++;
        if (i >= n) break;
        int a = 0, b = 0;
        if (!number(a)) return {};
        b = a;
        if (i < n && s[i] == '-') { i++; if (!number(b) || b < a) return {}; }
        for (int c = a; c <= b; c++) out.push_back(c);
        while (i < n && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t')) i++;
        if (i < n) { if (s[i] != ',') return {}; i++; }
    }
    return out;
}

static bool numa_enabled() {
    const char* v = getenv("MP2V_NUMA");
    return !(v && atoi(v) == 0);
}

static int count_numa_nodes() {
    DIR* d = opendir("/sys/devices/system/node");
    if (!d) return 0;
    int nodes = 0;
    while (dirent* e = readdir(d))
        if (!strncmp(e->d_name, "node", 4) && isdigit((unsigned char)e->d_name[4])) nodes++;
    closedir(d);
    return nodes;
}

int numa_node_of_pci_device(const std::string& bus_id) {
    if (!numa_enabled() || bus_id.empty() || count_numa_nodes() < 2) return -1;
    std::string id = bus_id;
    std::transform(id.begin(), id.end(), id.begin(), [](unsigned char c) { return (char)tolower(c); });
    std::ifstream f("/sys/bus/pci/devices/" + id + "/numa_node");
    int node = -1;
    if (!(f >> node)) return -1;
    return node;
}

std::vector<int> cpus_of_numa_node(int node) {
    if (node < 0) return {};
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string s;
    if (!std::getline(f, s)) return {};
    return parse_cpu_list(s);
}

static bool set_affinity_to(const std::vector<int>& cpus) {
    if (cpus.empty()) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    bool any = false;
    for (int c : cpus) if (c >= 0 && c < CPU_SETSIZE) { CPU_SET(c, &set); any = true; }
    return any && sched_setaffinity(0, sizeof(set), &set) == 0;
}

numa_scope_t::numa_scope_t(int node) {
    if (node < 0) return;
    cpu_set_t prev;
    CPU_ZERO(&prev);
    if (sched_getaffinity(0, sizeof(prev), &prev) != 0) return;
    if (!set_affinity_to(cpus_of_numa_node(node))) return;
    saved_.resize(sizeof(prev) / sizeof(unsigned long));
    memcpy(saved_.data(), &prev, sizeof(prev));
    bound_ = true;
}

numa_scope_t::~numa_scope_t() {
    if (!bound_) return;
    cpu_set_t prev;
    memcpy(&prev, saved_.data(), sizeof(prev));
    sched_setaffinity(0, sizeof(prev), &prev);
}

bool bind_this_thread_to_numa_node(int node) { return node >= 0 && set_affinity_to(cpus_of_numa_node(node)); }

}  // namespace mp2v
