/* oracle/slamshim: stand-in for DBoW2::FeatureVector (a map node id -> feature indices); only its type is needed to compile ORBmatcher.cc */
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
}
