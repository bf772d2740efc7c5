// TEST INFRASTRUCTURE shim for boost::disjoint_sets (boost is absent): union by rank with path
// compression, same interface as used by src/Align4.cpp:813-858.
#pragma once
namespace boost {
template<class RankPA, class ParentPA> class disjoint_sets {
public:
    disjoint_sets(RankPA r, ParentPA p) : rank(r), parent(p) {}
    template<class T> void make_set(T x) { parent[x] = x; rank[x] = 0; }
    template<class T> T find_set(T x)
    {
        T root = x;
        while(parent[root] != root) root = parent[root];
        while(parent[x] != root) { T next = parent[x]; parent[x] = root; x = next; }
        return root;
    }
    template<class T> void union_set(T x, T y)
    {
        x = find_set(x); y = find_set(y);
        if(x == y) return;
        if(rank[x] > rank[y]) parent[y] = x;
        else { parent[x] = y; if(rank[x] == rank[y]) ++rank[y]; }
    }
private:
    RankPA rank;
    ParentPA parent;
};
}
