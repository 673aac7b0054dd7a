int sforacle_placeholder(void){return 0;}
