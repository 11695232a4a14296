#include "examples.hpp"

namespace portrayer {
namespace {
std::map<std::string, ExampleFn>& mutable_registry() {
    static std::map<std::string, ExampleFn> r;
    return r;
}
}  // namespace
ExampleRegistrar::ExampleRegistrar(const std::string& name, ExampleFn fn) { mutable_registry()[name] = std::move(fn); }
const std::map<std::string, ExampleFn>& example_registry() { return mutable_registry(); }
}  // namespace portrayer
