// Stand-in for mpc_planner_util/parameters.h: the CONFIG singleton over settings.yaml (parameters.h:11,23-52).
#pragma once
#include <mpc_planner_util/load_yaml.hpp>
#include <ros_tools/logging.h>
#define LOG_MARK(x) do { } while (0)
#define CONFIG Configuration::getInstance().getYAMLNode()
class Configuration {
public:
    static Configuration& getInstance() { static Configuration instance; return instance; }
    void initialize(const std::string& config_file) { loadConfigYaml(config_file, _config); }
    YAML::Node& getYAMLNode() { return _config; }
private:
    YAML::Node _config;
    Configuration() {}
    Configuration(const Configuration&) = delete;
    Configuration& operator=(const Configuration&) = delete;
};
