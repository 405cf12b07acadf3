// TEST INFRASTRUCTURE (oracle shim) -- not product code.
// Stand-in for opm-core's ParameterGroup (third party, absent). Only what
// EulerUpstream_impl.hpp:95-108 needs: getDefault<T>(name, default), get<T>, has.
#ifndef ORACLE_SHIM_PARAMETERGROUP_HPP
#define ORACLE_SHIM_PARAMETERGROUP_HPP
#include <map>
#include <string>
#include <sstream>
#include <stdexcept>
namespace Opm { namespace parameter {
    class ParameterGroup {
    public:
        ParameterGroup() {}
        ParameterGroup(int argc, char** argv, bool = true)
        {
            for (int i = 1; i < argc; ++i) {
                std::string a(argv[i]);
                std::string::size_type eq = a.find('=');
                if (eq != std::string::npos) { kv_[a.substr(0, eq)] = a.substr(eq + 1); }
            }
        }
        template <typename T> void insertParameter(const std::string& name, const T& v)
        {
            std::ostringstream os; os.precision(17); os << std::boolalpha << v; kv_[name] = os.str();
        }
        bool has(const std::string& name) const { return kv_.count(name) != 0; }
        template <typename T> T get(const std::string& name) const
        {
            std::map<std::string, std::string>::const_iterator it = kv_.find(name);
            if (it == kv_.end()) { throw std::runtime_error("Missing parameter " + name); }
            return conv<T>(it->second);
        }
        template <typename T> T getDefault(const std::string& name, const T& d) const
        {
            return has(name) ? get<T>(name) : d;
        }
        bool anyUnused() const { return false; }
        void displayUsage() const {}
    private:
        template <typename T> static T conv(const std::string& s)
        {
            std::istringstream is(s); T v; is >> v; return v;
        }
        std::map<std::string, std::string> kv_;
    };
    template <> inline bool ParameterGroup::conv<bool>(const std::string& s)
    {
        return s == "true" || s == "1" || s == "True";
    }
    template <> inline std::string ParameterGroup::conv<std::string>(const std::string& s) { return s; }
}}
#endif
