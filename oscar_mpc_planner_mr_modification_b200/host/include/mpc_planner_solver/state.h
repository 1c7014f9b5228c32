// MPCPlanner::State -- same public surface as the reference's mpc_planner_solver/state.h:13-33.
#ifndef STATE_H
#define STATE_H

#include <mpc_planner_util/load_yaml.hpp>

#include <Eigen/Dense>

#include <string>
#include <vector>

namespace MPCPlanner
{
    struct State
    {
        State();

        void initialize();

        double get(std::string &&var_name) const;
        Eigen::Vector2d getPos() const;

        void set(std::string &&var_name, double value);
        void print() const;

        bool validData() const;

    private:
        std::vector<double> _state;
        YAML::Node _config, _model_map;

        int _nu;
    };
}

#endif // STATE_H
