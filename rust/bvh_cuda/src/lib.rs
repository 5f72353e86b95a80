//! Drop-in for `crates/bvh` (same public names and signatures: crates/bvh/src/lib.rs:5-7) that forwards to the
//! sm_100a CUDA library through the C ABI of include/bvh_cuda.h.  Swap `bvh = { path = "crates/bvh" }` for
//! `bvh = { package = "bvh_cuda", path = "rust/bvh_cuda" }` in crates/pools/Cargo.toml and the root Cargo.toml.
//!
//! NOT compiled in this repository's image (no cargo/rustc); kept in sync with the header by hand.
use bytemuck::{Pod, Zeroable};
use components::{Instance, MeshInfo};
use glam::{UVec3, Vec3};
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone, Default, Debug, Pod, Zeroable)]
pub struct BvhNode {
    pub min: Vec3,
    pub left_first: u32,
    pub max: Vec3,
    pub count: u32,
}

#[repr(C)]
#[derive(Copy, Clone, Default, Debug, Pod, Zeroable)]
pub struct TlasNode {
    pub min: Vec3,
    pub left_right: u32,
    pub max: Vec3,
    pub instance_idx: u32,
}

impl TlasNode {
    pub fn is_leaf(&self) -> bool {
        self.left_right == 0
    }
}

#[derive(PartialOrd, PartialEq, Clone, Copy, Debug)]
pub enum Dist {
    Hit(f32),
    Miss,
}

#[derive(Clone, Copy, Default, Debug)]
pub struct Ray {
    pub orig: Vec3,
    pub dir: Vec3,
}

impl Ray {
    pub fn new(orig: Vec3, dir: Vec3) -> Self {
        Self { orig, dir }
    }
}

#[allow(non_camel_case_types)]
type ctx_t = c_void;

extern "C" {
    fn bvh_cuda_create(device: c_int, out: *mut *mut ctx_t) -> c_int;
    fn bvh_cuda_last_error(ctx: *const ctx_t) -> *const c_char;
    fn bvh_cuda_blas_build(
        ctx: *mut ctx_t, vertices: *const f32, n_vertices: usize, indices: *mut u32, n_tris: usize,
        nodes_out: *mut BvhNode, nodes_cap: usize, n_nodes_out: *mut u32,
    ) -> c_int;
    fn bvh_cuda_tlas_build(
        ctx: *mut ctx_t, instances: *const Instance, n_inst: usize, meshes: *const MeshInfo, n_mesh: usize,
        nodes_out: *mut TlasNode, children_out: *mut u32,
    ) -> c_int;
    fn bvh_cuda_trace_blas(
        ctx: *mut ctx_t, nodes: *const BvhNode, n_nodes: usize, vertices: *const f32, n_vertices: usize,
        indices: *const u32, n_tris: usize, ray_o: *const f32, ray_d: *const f32, n_rays: usize,
        t_out: *mut f32, tri_out: *mut u32,
    ) -> c_int;
}

thread_local! {
    // one context per host thread (the reference path is single-threaded: crates/app/src/lib.rs:110-114)
    static CTX: *mut ctx_t = unsafe {
        let mut p: *mut ctx_t = std::ptr::null_mut();
        let rc = bvh_cuda_create(0, &mut p);
        assert!(rc == 0, "bvh_cuda_create failed ({rc}): a CUDA device is required, there is no CPU fallback");
        p
    };
}

fn check(ctx: *mut ctx_t, rc: c_int) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(bvh_cuda_last_error(ctx)) };
        // the reference panics on the same inputs (blas.rs:84, mesh/mod.rs:321) or never returns (blas.rs:115,139)
        panic!("bvh_cuda error {rc}: {}", msg.to_string_lossy());
    }
}

pub struct BvhBuilder<'a> {
    num_bins: usize,
    vertices: &'a [Vec3],
    indices: &'a mut [UVec3],
}

impl<'a> BvhBuilder<'a> {
    pub fn new(vertices: &'a [Vec3], indices: &'a mut [UVec3]) -> Self {
        Self { num_bins: 8, vertices, indices }
    }

    /// Stored and never read, exactly like crates/bvh/src/blas.rs:64-67 / :136.
    pub fn set_bin_number(mut self, num_bins: usize) -> Self {
        self.num_bins = num_bins;
        self
    }

    pub fn build(self) -> Bvh {
        let n = self.indices.len();
        let mut nodes = vec![BvhNode::default(); n * 2];
        let mut used = 0u32;
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_blas_build(
                ctx, self.vertices.as_ptr() as *const f32, self.vertices.len(),
                self.indices.as_mut_ptr() as *mut u32, n, nodes.as_mut_ptr(), nodes.len(), &mut used,
            );
            check(ctx, rc);
        });
        nodes.truncate(used as usize);
        Bvh { nodes }
    }
}

pub struct Bvh {
    pub nodes: Vec<BvhNode>,
}

impl Bvh {
    pub fn traverse_iter(&self, vertices: &[Vec3], indices: &[UVec3], ray: Ray) -> Dist {
        let (mut t, mut tri) = (0f32, 0u32);
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_trace_blas(
                ctx, self.nodes.as_ptr(), self.nodes.len(), vertices.as_ptr() as *const f32, vertices.len(),
                indices.as_ptr() as *const u32, indices.len(), &ray.orig as *const Vec3 as *const f32,
                &ray.dir as *const Vec3 as *const f32, 1, &mut t, &mut tri,
            );
            check(ctx, rc);
        });
        if tri == u32::MAX { Dist::Miss } else { Dist::Hit(t) }
    }
}

pub struct Tlas {
    pub nodes: Vec<TlasNode>,
    /// unpacked child ids per node (needed by a traversal above 32 767 instances; left_right packs 16+16 bits)
    pub children: Vec<[u32; 2]>,
}

impl Tlas {
    pub fn empty() -> Self {
        Self { nodes: vec![], children: vec![] }
    }

    pub fn build(&mut self, instances: &[Instance], meshes: &[MeshInfo]) {
        let total = 2 * instances.len() + 1;
        self.nodes = vec![TlasNode::default(); total];
        self.children = vec![[0u32; 2]; total];
        CTX.with(|&ctx| unsafe {
            let rc = bvh_cuda_tlas_build(
                ctx, instances.as_ptr(), instances.len(), meshes.as_ptr(), meshes.len(),
                self.nodes.as_mut_ptr(), self.children.as_mut_ptr() as *mut u32,
            );
            check(ctx, rc);
        });
    }
}
