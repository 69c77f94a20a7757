// TEST INFRASTRUCTURE ONLY (oracle build shim): the few tbb::flow names graph.h
// needs to declare its stage functors.  Ports just record what is put.
#pragma once
#include <tuple>
#include <vector>
#include <atomic>
#include <deque>
#include <string>
#include <cstring>
#include <cstdio>
namespace tbb { namespace flow {
using std::tuple; using std::get;
template <typename T> struct shim_port {
    std::vector<T> items;
    bool try_put(const T& t) { items.push_back(t); return true; }
};
template <typename In, typename Out> struct multifunction_node;
template <typename In, typename... O> struct multifunction_node<In, std::tuple<O...>> {
    typedef std::tuple<shim_port<O>...> output_ports_type;
};
} }
