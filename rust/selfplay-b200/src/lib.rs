//! Safe wrapper with the signatures `learner.rs` / `learner_concurrent.rs` use today
//! (ref: src/mcts.rs:32-44, :86, :161, :196).  The trees live on the GPU; a `Tree` is a handle on one
//! engine slot.  NOT compiled in the build image (no rustc/cargo there).
use selfplay_b200_sys as sys;
use std::ffi::CStr;
use std::marker::PhantomData;

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }
pub type Result<T> = std::result::Result<T, Error>;

/// Conversion between the reference's game states and the ABI's bitboard state.
/// Implemented in the reference crate for `game::connect_four::State` (bit = col*7+row) and
/// `game::tictactoe::State` (bit = row*3+col); see INTEGRATION.md.
pub trait AbiState: Sized {
    const GAME: i32;
    const NUM_ACTIONS: usize;
    fn to_abi(&self) -> sys::spb_state;
    fn from_abi(s: &sys::spb_state) -> Self;
}

/// ref: `Args` mcts.rs:8-18 — only `c` and `num_searches` are read on the search path.
#[derive(Clone, Copy)]
pub struct Args { pub c: f32, pub num_searches: u32, pub num_parallel_self_play_games: usize }
impl Default for Args {
    fn default() -> Self { Args { c: 2.0, num_searches: 600, num_parallel_self_play_games: 100 } }
}

/// ref: `Mcts<T>` mcts.rs:41-44.  Owns the engine (node pools, weights) of one GPU; one per host thread.
pub struct Mcts<S: AbiState> { raw: *mut sys::spb_engine, pub args: Args, _s: PhantomData<S> }
unsafe impl<S: AbiState> Send for Mcts<S> {}

/// ref: `Tree<T>` mcts.rs:32-39: a handle on slot `slot`; histories stay on the host like in the reference.
pub struct Tree<S: AbiState> { pub slot: u32, pub state_history: Vec<S>, pub policy_history: Vec<Vec<f32>> }

impl<S: AbiState> Mcts<S> {
    pub fn new(args: Args, device: i32, evaluator: i32) -> Result<Self> {
        let mut cfg = sys::spb_config::default();
        unsafe { sys::spb_default_config(&mut cfg) };
        cfg.game = S::GAME; cfg.device = device; cfg.c = args.c; cfg.evaluator = evaluator;
        cfg.num_games = args.num_parallel_self_play_games as u32;
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { sys::spb_create(&cfg, &mut raw) };
        if rc != sys::SPB_OK { return Err(Error { code: rc, message: last_error(std::ptr::null()) }); }
        Ok(Mcts { raw, args, _s: PhantomData })
    }
    fn check(&self, rc: i32) -> Result<()> {
        if rc == sys::SPB_OK { Ok(()) } else { Err(Error { code: rc, message: last_error(self.raw) }) }
    }
    /// ref: `VarStore::load` main.rs:61 — bytes of the safetensors file `var_store.save` wrote (learner.rs:192).
    pub fn load_weights(&mut self, safetensors: &[u8]) -> Result<()> {
        self.check(unsafe { sys::spb_load_weights(self.raw, safetensors.as_ptr() as *const _, safetensors.len()) })
    }
    /// ref: `Tree::with_root_state` mcts.rs:86 / `Tree::default` mcts.rs:67.
    pub fn new_tree(&mut self, slot: u32, root: Option<&S>) -> Result<Tree<S>> {
        let abi = root.map(|s| s.to_abi());
        let p = abi.as_ref().map_or(std::ptr::null(), |s| s as *const _);
        self.check(unsafe { sys::spb_reset_games(self.raw, &slot, 1, p) })?;
        Ok(Tree { slot, state_history: Vec::new(), policy_history: Vec::new() })
    }
    /// ref: `Mcts::search` mcts.rs:196 — result i belongs to trees[i]:
    /// (root visit counts scattered by action and normalised, [(child arena id, visit count as f32)] in child order).
    pub fn search(&mut self, trees: &mut Vec<&mut Tree<S>>) -> Result<Vec<(Vec<f32>, Vec<(usize, f32)>)>> {
        self.check(unsafe { sys::spb_search(self.raw, self.args.num_searches) })?;
        let mut out = Vec::with_capacity(trees.len());
        for t in trees.iter() {
            let (mut a, mut c, mut ids, mut n) = ([0u8; sys::SPB_MAX_ACTIONS], [0u32; sys::SPB_MAX_ACTIONS], [0u32; sys::SPB_MAX_ACTIONS], 0u32);
            self.check(unsafe { sys::spb_root_children(self.raw, t.slot, a.as_mut_ptr(), c.as_mut_ptr(), ids.as_mut_ptr(), &mut n) })?;
            let mut policy = vec![0f32; S::NUM_ACTIONS];
            self.check(unsafe { sys::spb_root_policy(self.raw, t.slot, policy.as_mut_ptr()) })?;
            out.push((policy, (0..n as usize).map(|i| (ids[i] as usize, c[i] as f32)).collect()));
        }
        Ok(out)
    }
    /// ref: `Tree::use_subtree` mcts.rs:161; returns the new root state (`tree.arena[0].state`).
    pub fn use_subtree(&mut self, tree: &mut Tree<S>, new_root_id: usize) -> Result<S> {
        let id = new_root_id as u32;
        let mut st = sys::spb_state::default();
        self.check(unsafe { sys::spb_advance(self.raw, &tree.slot, &id, 1, &mut st) })?;
        Ok(S::from_abi(&st))
    }
    /// ref: `tree.arena[id].state` (learner_concurrent.rs:184,195; main.rs:92).
    pub fn node_state(&mut self, tree: &Tree<S>, node_id: usize) -> Result<S> {
        let mut st = sys::spb_state::default();
        self.check(unsafe { sys::spb_get_state(self.raw, tree.slot, node_id as u32, &mut st) })?;
        Ok(S::from_abi(&st))
    }
}

impl<S: AbiState> Drop for Mcts<S> {
    fn drop(&mut self) { unsafe { sys::spb_destroy(self.raw); } }
}

fn last_error(e: *const sys::spb_engine) -> String {
    unsafe { CStr::from_ptr(sys::spb_last_error(e)).to_string_lossy().into_owned() }
}
