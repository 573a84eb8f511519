// Stand-in for Thirdparty/DBoW2/DBoW2/FeatureVector.h: node id -> indices of the features that fall under that vocabulary node.
// Test scaffolding only; a real build includes the reference's own header.
#pragma once
#include <map>
#include <vector>
namespace DBoW2
{
typedef unsigned int NodeId;
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>>
{
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2
