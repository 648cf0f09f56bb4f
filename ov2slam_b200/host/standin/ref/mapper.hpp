// The part of the reference's include/mapper.hpp (:75-133) the matchToMap drop-in needs: the class, its state pointers and the
// one method.  Stand-in for the container's compile check; a real build includes the reference's own header.
#pragma once

#include <map>
#include <unordered_set>

#include "map_manager.hpp"

class Mapper {
public:
    Mapper() {}
    Mapper(std::shared_ptr<SlamParams> pslamstate, std::shared_ptr<MapManager> pmap) : pslamstate_(pslamstate), pmap_(pmap) {}

    std::map<int,int> matchToMap(const Frame &frame, const float fmaxprojerr, const float fdistratio, std::unordered_set<int> &set_local_lmids);

    std::shared_ptr<SlamParams> pslamstate_;
    std::shared_ptr<MapManager> pmap_;

    bool bnewkfavailable_ = false;
};
