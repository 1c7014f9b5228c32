// Stand-in for mpc_planner_util/load_yaml.hpp.  The reference resolves the three generated maps as
// <source dir>/../config/<name>.yaml (load_yaml.hpp:10-11); outside a catkin workspace the emitter's
// output directory is given by $MPCGPU_CONFIG_DIR (or the compile-time default).
#pragma once
#include <yaml-cpp/yaml.h>

#include <cstdlib>
#include <string>
#ifndef MPCGPU_DEFAULT_CONFIG_DIR
#define MPCGPU_DEFAULT_CONFIG_DIR "."
#endif
inline std::string mpcgpu_config_dir()
{
    const char* e = std::getenv("MPCGPU_CONFIG_DIR");
    return e ? std::string(e) : std::string(MPCGPU_DEFAULT_CONFIG_DIR);
}
#define SYSTEM_CONFIG_PATH(x, filename) (mpcgpu_config_dir() + "/" + filename + ".yaml")
inline void loadConfigYaml(const std::string& file, YAML::Node& _yaml_out) { _yaml_out = YAML::LoadFile(file); }
